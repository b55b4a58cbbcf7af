# e2e of the host mirror: chunks, huge pages
mkdir -p gpurun_out
cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag 2>/dev/null
q() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
e=d['e2e']
print('$1', 'value', round(d['value']), 'e2e', round(e['value']), 'k_step_ms_in_e2e', round(e['k_step_ms_in_e2e'],4), e['mirror_last_step_us'], 'ok', e['mirror_equals_device'])"; }
for k in 16 32; do AGARCL_MIRROR_CHUNKS=$k python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>>gpurun_out/exp.err | q chunks$k; done
AGARCL_MIRROR_NO_THP=1 AGARCL_MIRROR_CHUNKS=32 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>>gpurun_out/exp.err | q nothp_chunks32
grep -i huge /proc/meminfo | head -3
(timeout 900 python -m pytest tests/test_gpu_mirror.py -x -q 2>&1 | tail -5)
