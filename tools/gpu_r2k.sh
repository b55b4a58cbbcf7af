for tb in 6 70 6 70; do echo "tick_barrier $tb"; AGARCL_TICK_BARRIER=$tb timeout 300 python tools/exp_perstep.py 2000 40 2>&1 | tail -1; done
AGARCL_TICK_BARRIER=70 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 2>&1 | tail -3
