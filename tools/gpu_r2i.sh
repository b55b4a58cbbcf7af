for w in 16 15 14 13 12; do echo "warps $w"; AGARCL_WARPS=$w python tools/exp_perstep.py 2000 40 2>&1 | tail -1; done
