# one GPU visit: parity suite, bench (both arms), ncu launch list, one full capture of k_step (steady-state games)
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log)
tail -5 gpurun_out/gpu_tests.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# launches before the timed region of `--steps 20 --warmup 3`: 100 + 3 + 20 + (2000 - 123) + 3 = 2003
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2003 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 2003 -c 2 -o gpurun_out/k_step_full -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'flags',d['config']['state_flags_seen'], 'age', d['age_profile'], 'cpu', d['cpu_baseline']['value'])
r=json.load(open('gpurun_out/bench_ref.json'));print('ref arm', r.get('value'), r.get('unavailable'))"
