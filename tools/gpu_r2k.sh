for pf in 0 1 0 1; do echo "prefetch_next $pf"; AGARCL_PREFETCH_NEXT=$pf timeout 300 python tools/exp_perstep.py 2000 40 2>&1 | tail -1; done
