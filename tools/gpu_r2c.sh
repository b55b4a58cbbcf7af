mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log); tail -3 gpurun_out/gpu_tests.log
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'young', d['age_profile'][0], 'int16', d['int16_profile'])"
AGARCL_AUTO_SCHEDULE=0 timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-int16 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('auto off: ms',d['ms_per_step'],'young', d['age_profile'][0])"
bash tools/gpu_sanitize.sh
