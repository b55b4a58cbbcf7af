# e2e of the host mirror against the number of chunks the batch is fetched in
mkdir -p gpurun_out
q() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
e=d['e2e']
print('$1', 'value', round(d['value']), 'e2e', round(e['value']), 'k_step_ms_in_e2e', round(e['k_step_ms_in_e2e'],4), e['mirror_last_step_us'], 'ok', e['mirror_equals_device'])"; }
for k in 8 16; do AGARCL_MIRROR_CHUNKS=$k python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>>gpurun_out/exp.err | q chunks$k; done
(timeout 900 python -m pytest tests/test_gpu_mirror.py -x -q 2>&1 | tail -5)
