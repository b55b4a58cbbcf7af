#!/usr/bin/env python
"""Per-step device time at steady state (configs[1]): every 10th tick is a bot-decision tick (Engine.hpp:498-499), i.e. two of
every five env-steps contain one.  usage: python tools/exp_perstep.py [settle] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from agarcl_b200 import make_cfg
from agarcl_b200.batch import Batch
import bench

settle = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
N = 4096
b = Batch(make_cfg(n_instances=N, **bench.WORKLOAD))
b.seed(np.arange(N, dtype=np.uint64) + np.uint64(1))
b.reset()
stream = torch.cuda.current_stream().cuda_stream
gen = torch.Generator(device="cuda"); gen.manual_seed(1234)
dxdy = (torch.rand((16, N, 2), device="cuda", generator=gen) * 2 - 1).float().contiguous()
act = torch.randint(0, 3, (16, N), device="cuda", generator=gen, dtype=torch.int32).contiguous()
def one(i):
    b.set_actions_device(dxdy[i % 16].data_ptr(), act[i % 16].data_ptr(), stream)
    b.step(stream)
for i in range(settle): one(i)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
ev[0].record()
for i in range(K):
    one(i); ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
tick0 = settle * 4
dec = [any((tick0 + 4 * i + t) % 10 == 0 for t in range(4)) for i in range(K)]
d = [m for m, f in zip(ms, dec) if f]; n = [m for m, f in zip(ms, dec) if not f]
print("per-step ms:", " ".join(f"{m:.2f}{'*' if f else ''}" for m, f in zip(ms, dec)))
print(f"steps with a decision tick: {np.mean(d):.3f} ms ({len(d)}), without: {np.mean(n):.3f} ms ({len(n)}), all: {np.mean(ms):.3f}")
