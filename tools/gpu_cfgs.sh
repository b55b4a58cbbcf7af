# every BASELINE config through bench.py (short runs), the tasks_configs runner, the vector-env tests
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -x -q --timeout 300 > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log); tail -3 gpurun_out/gpu_tests.log
for c in c1 c3 c4 c5; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -2 gpurun_out/bench_$c.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_$c.json'));print('$c','value',round(d['value']),'ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],4),d['roofline']['kernel'],'e2e',round(d['e2e']['value']),'flags',d['config']['state_flag_instances'],'cpu',d['cpu_baseline'] and round(d['cpu_baseline']['value'] or 0), 'mirror', d['e2e'].get('mirror_equals_device'))"
done
timeout 600 python tools/run_tasks_configs.py --envs 4096 --steps 300 > gpurun_out/tasks_configs.jsonl 2> gpurun_out/tasks.err; tail -2 gpurun_out/tasks.err; cat gpurun_out/tasks_configs.jsonl | cut -c1-250
