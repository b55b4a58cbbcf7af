python tools/exp_perstep.py 2000 40 2>&1 | tail -1
python tools/exp_perstep.py 2000 40 2>&1 | tail -1
