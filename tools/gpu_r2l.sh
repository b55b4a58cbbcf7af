mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 110 -c 1 -o gpurun_out/k_step_young -f python tools/exp_perstep.py 100 20 > gpurun_out/ncu_young.log 2>&1
tail -2 gpurun_out/ncu_young.log | cut -c1-200
