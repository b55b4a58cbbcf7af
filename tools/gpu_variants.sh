# times every library under variants/ on steady-state games (the box's copy of the tree is scratch: the product library is overwritten there)
for v in variants/*.so; do
  cp $v agarcl_b200/libagarcl_b200.so
  echo "== $v"
  timeout 300 python tools/exp_perstep.py 2000 40 2>&1 | tail -1
  timeout 300 python tools/exp_perstep.py 2000 40 2>&1 | tail -1
done
