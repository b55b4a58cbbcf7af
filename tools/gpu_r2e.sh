mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -k "ram or gobigger or vector_env" > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log); tail -3 gpurun_out/gpu_tests.log
timeout 900 python bench.py --config c3 --steps 30 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err
python -c "
import json;d=json.load(open('gpurun_out/bench_c3.json'));print('c3 value',round(d['value']),'ms',round(d['ms_per_step'],4),[(k['kernel'], round(k['avg_launch_ms'],4), round(k['frac'],4)) for k in d['roofline_all']['kernels']],'e2e',round(d['e2e']['value']),'cpu',d['cpu_baseline'] and round(d['cpu_baseline']['value'] or 0))"
