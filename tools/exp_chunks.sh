mkdir -p gpurun_out
q() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],4), [ (k['kernel'], round(k['avg_launch_ms'],4)) for k in d['roofline_all']['kernels']])"; }
for z in 0 1 2 4 8 12 16; do AGARCL_ZERO_CHUNKS=$z python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>>gpurun_out/exp.err | q chunks$z; done
