# one GPU visit: the whole -m gpu suite, then the per-step device time of steady-state games twice (tools/exp_perstep.py)
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log); tail -8 gpurun_out/gpu_tests.log
python tools/exp_perstep.py 2000 40 2>&1 | tail -1
python tools/exp_perstep.py 2000 40 2>&1 | tail -1
