mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 500 > gpurun_out/gpu_tests_multi.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests_multi.log); tail -5 gpurun_out/gpu_tests_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-int16 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -2 gpurun_out/bench_n2.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]);print('n2 value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']), d['e2e']['mirror_last_step_us'], 'mirror ok', d['e2e']['mirror_equals_device'])"
nproc
