# one GPU visit: parity suite, bench (both arms), the other BASELINE configs, ncu launch list, one full capture of k_step (steady-state games)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log)
tail -5 gpurun_out/gpu_tests.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
for c in c1 c3 c4 c5; do
  timeout 900 python bench.py --config $c --steps 30 --warmup 5 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
done
timeout 600 python tools/run_tasks_configs.py --envs 4096 --steps 600 > gpurun_out/tasks_configs.jsonl 2> gpurun_out/tasks.err
# launches before the timed region of `--steps 20 --warmup 3`: k_step + k_order per step -> 2 * (100 + 3 + 20 + (2000 - 123) + 3) = 4006
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4006 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-int16 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 2003 -c 2 -o gpurun_out/k_step_full -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-int16 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'flags',d['config']['state_flag_instances'], 'age', d['age_profile'], 'cpu', d['cpu_baseline']['value'], 'int16', d['int16_profile'])
r=json.load(open('gpurun_out/bench_ref.json'));print('ref arm', r.get('value'), r.get('unavailable'))"
