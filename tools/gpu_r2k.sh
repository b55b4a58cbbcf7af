for w in 16 14 15 16 14; do echo "warps $w"; AGARCL_WARPS=$w timeout 300 python tools/exp_perstep.py 2000 40 2>&1 | tail -1; done
