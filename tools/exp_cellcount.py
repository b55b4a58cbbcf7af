#!/usr/bin/env python
"""Distribution of live cells per instance at steady state (configs[1], age 2000): players_collision stages the cells of an instance in
shared memory up to kSnapCap."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from agarcl_b200 import make_cfg
from agarcl_b200.batch import Batch
import bench

N = 4096
b = Batch(make_cfg(n_instances=N, **bench.WORKLOAD))
b.seed(np.arange(N, dtype=np.uint64) + np.uint64(1))
b.reset()
stream = torch.cuda.current_stream().cuda_stream
gen = torch.Generator(device="cuda"); gen.manual_seed(1234)
dxdy = (torch.rand((16, N, 2), device="cuda", generator=gen) * 2 - 1).float().contiguous()
act = torch.randint(0, 3, (16, N), device="cuda", generator=gen, dtype=torch.int32).contiguous()
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2000):
    b.set_actions_device(dxdy[i % 16].data_ptr(), act[i % 16].data_ptr(), stream)
    b.step(stream)
torch.cuda.synchronize()
tot, multi, foods = [], [], []
for i in range(0, N, 4):
    sv = b.download_state(i)
    n = sv.players["n_cells"]
    tot.append(int(n.sum())); multi.append(int((n >= 2).sum())); foods.append(int(sv.hdr["n_foods"]))
tot = np.array(tot)
print("instances sampled", len(tot), "cells per instance: mean", tot.mean(), "p50", np.percentile(tot, 50), "p90", np.percentile(tot, 90),
      "p99", np.percentile(tot, 99), "max", tot.max(), " > 96:", float((tot > 96).mean()), " > 80:", float((tot > 80).mean()))
print("multi-cell players per instance: mean", np.mean(multi), "max", np.max(multi), " foods: mean", np.mean(foods), "max", np.max(foods))
