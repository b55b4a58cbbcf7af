mkdir -p gpurun_out
nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus ${NG:-8} --steps 20 --warmup 5 --no-int16 > gpurun_out/bench_n${NG:-8}.json 2> gpurun_out/bench_n${NG:-8}.err; tail -2 gpurun_out/bench_n${NG:-8}.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_n${NG:-8}.json').read().strip().splitlines()[-1]);print('n8: value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']), 'lists', round(d['e2e']['lists']['value']), d['e2e']['mirror_last_step_us'], 'mirror ok', d['e2e']['mirror_equals_device'], d['e2e']['lists']['decodes_to_device_observation'])"
