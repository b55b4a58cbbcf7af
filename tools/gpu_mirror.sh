# GPU check of the host mirror: its parity tests, then one bench line (no CPU baseline)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_mirror.py -x -q > gpurun_out/mirror_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/mirror_tests.log)
tail -25 gpurun_out/mirror_tests.log
nproc; free -g | head -2
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e'],'flags',d['config']['state_flags_seen'])"
