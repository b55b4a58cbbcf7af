mkdir -p gpurun_out
q() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],4), [ (k['kernel'], round(k['avg_launch_ms'],4)) for k in d['roofline_all']['kernels']])"; }
echo start; AGARCL_FUSE_CLEAR=0 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>>gpurun_out/exp.err | q fuse0
AGARCL_FUSE_CLEAR=1 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>>gpurun_out/exp.err | q fuse1
for t in 0 1 2 4 8; do python bench.py --steps 40 --warmup 5 --no-cpu-baseline --tps $t 2>>gpurun_out/exp.err | q tps$t; done
