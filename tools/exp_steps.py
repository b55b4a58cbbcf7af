"""Per-step device times of the bench workload at steady state (diagnostic): one CUDA event per step."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from agarcl_b200 import make_cfg
from agarcl_b200.batch import Batch
N = 4096
settle = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
cfg = make_cfg(n_instances=N, device=0, **bench.WORKLOAD)
b = Batch(cfg); b.seed(np.arange(N, dtype=np.uint64) + np.uint64(1)); b.reset()
gen = torch.Generator(device="cuda"); gen.manual_seed(1234)
dxdy = (torch.rand((16, N, 2), device="cuda", generator=gen) * 2 - 1).float().contiguous()
act = torch.randint(0, 3, (16, N), device="cuda", generator=gen, dtype=torch.int32).contiguous()
s = torch.cuda.current_stream().cuda_stream
def step(i):
    b.set_actions_device(dxdy[i % 16].data_ptr(), act[i % 16].data_ptr(), s); b.step(s)
for i in range(settle): step(i)
torch.cuda.synchronize()
K = 40
ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
ev[0].record()
for i in range(K):
    step(i); ev[i + 1].record()
torch.cuda.synchronize()
print("back-to-back ms:", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(K)])
ts = []
for i in range(K):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); step(i); e1.record(); torch.cuda.synchronize(); ts.append(round(e0.elapsed_time(e1), 3))
print("isolated ms:", ts)
b.close()
