mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log); tail -4 gpurun_out/gpu_tests.log
python tools/exp_perstep.py 2000 40 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 2003 -c 1 -o gpurun_out/k_step_r2j -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-int16 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
