for tb in 6 2 4 7 14 22; do echo "tick_barrier $tb"; AGARCL_TICK_BARRIER=$tb timeout 300 python tools/exp_perstep.py 2000 40 2>&1 | tail -1; done
