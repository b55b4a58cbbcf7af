# after tools/gpu_round.sh: turn gpurun_out/ into the committed summaries under profiles/ (tag = $1, e.g. r01m)
tag=${1:-r01m}
python profiles/summarize_ncu.py gpurun_out/k_step_full.ncu-rep > profiles/${tag}_k_step_steady_ncu_full.txt
cp gpurun_out/launches.csv profiles/${tag}_launches.csv
cp gpurun_out/bench.json profiles/${tag}_bench.json
cp gpurun_out/bench_ref.json profiles/${tag}_bench_reference_arm.json
(cd gpurun_out && ncu -i k_step_full.ncu-rep --page source --csv --print-source cuda,sass > src_final.csv 2>/dev/null; python ../tools/ncu_phases.py src_final.csv 10 > ../profiles/${tag}_k_step_samples_by_function.txt 2>&1)
python - <<PY
import json,csv,subprocess
out=subprocess.run(['ncu','-i','gpurun_out/k_step_full.ncu-rep','--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines())); h=rows[0]; u=rows[1]
f={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}
rd=[float(r[h.index('dram__bytes_read.sum')])*f[u[h.index('dram__bytes_read.sum')]] for r in rows[2:]]
wr=[float(r[h.index('dram__bytes_write.sum')])*f[u[h.index('dram__bytes_write.sum')]] for r in rows[2:]]
t={"k_step": int(sum(rd)/len(rd)+sum(wr)/len(wr)),
   "_source": "profiles/${tag}_k_step_steady_ncu_full.txt: dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the launches captured by 'ncu --set full -k regex:k_step -s 2003 -c 2 python bench.py --steps 4 --warmup 3' (steady-state games, tools/gpu_round.sh)",
   "_read": int(sum(rd)/len(rd)), "_write": int(sum(wr)/len(wr))}
json.dump(t,open('profiles/traffic.json','w'),indent=1); print(t['k_step'])
PY
