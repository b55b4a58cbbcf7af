# GPU check used during development: parity suite + one bench line (no CPU baseline)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 --durations=8 > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log)
tail -25 gpurun_out/gpu_tests.log
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'flags',d['config']['state_flags_seen'], 'young', d['age_profile'][0])"
