#!/usr/bin/env python
"""Per-instance cost distribution at steady state (configs[1]) and what it allows: a round of a CTA costs about its most expensive
instance, so the makespan of a launch is bounded below by how the sorted costs can be packed into rounds of 16 on 148 SMs."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from agarcl_b200 import make_cfg, _lib
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
prev = None
for i in range(2006):
    b.set_actions_device(dxdy[i % 16].data_ptr(), act[i % 16].data_ptr(), stream)
    b.step(stream)
    if i >= 2000:
        c = np.zeros(N, np.uint32)
        _lib.check(_lib.lib().agarcl_batch_costs(b._h, C.c_void_p(stream), c.ctypes.data_as(C.c_void_p)))
        c = c.astype(np.float64) / 1e6
        s = np.sort(c)[::-1]
        line = f"step {i}: Mcycles max {s[0]:.2f} p99 {s[40]:.2f} p90 {s[409]:.2f} p58 (rank 2368) {s[2368]:.2f} p50 {s[2048]:.2f} p10 {s[3686]:.2f} min {s[-1]:.2f} mean {c.mean():.2f}"
        # stripes of 16 in sorted order: a stripe costs its first (largest) entry; pairing bound: 148 CTAs, 256 stripes
        st = s[::16]
        pair = max(st[0], max(st[108 + k] + st[255 - k] for k in range(0, 74)))  # 40 alone... (st[0..39]) rest paired extremes
        line += f" | stripes: top {st[0]:.2f} #40 {st[40]:.2f} #147 {st[147]:.2f} #255 {st[255]:.2f}; sum all stripes/148 = {st.sum()/148:.2f}"
        if prev is not None:
            line += f" | corr with previous step {np.corrcoef(prev, c)[0,1]:.3f}, mean |rel change| {np.mean(np.abs(c-prev)/np.maximum(prev,1e-9)):.3f}"
        print(line)
        prev = c
