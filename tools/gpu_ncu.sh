# one full ncu capture of k_step at steady state (launch index: 2000 settle + warmups); $1 = extra skip
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 2003 -c 1 -o gpurun_out/k_step_full -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
