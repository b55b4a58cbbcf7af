for v in base l1big base l1big; do
  cp variants/$v.so agarcl_b200/libagarcl_b200.so
  echo "== $v, 14 warps"; AGARCL_WARPS=14 timeout 300 python tools/exp_perstep.py 2000 40 2>&1 | grep -v per-step | tail -2
done
