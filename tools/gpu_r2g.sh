mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_mirror.py tests/test_gpu_api.py -m gpu -x -q --timeout 600 > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log); tail -12 gpurun_out/gpu_tests.log
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-int16 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'lists',d['e2e']['lists'], 'mirror', d['e2e']['mirror_equals_device'])"
