# two GPUs: the NCCL tests, then the bench at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 300 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-int16 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -1 gpurun_out/bench_n2.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]);print('n2: value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']), 'lists', round(d['e2e']['lists']['value']))"
