# ncu capture of k_step at the bench's game stage: launch list + two full captures
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 104 -c 2 -o gpurun_out/k_step_full -f python bench.py --steps 4 --warmup 3 --settle 100 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
