#!/usr/bin/env python
"""Launches every kernel of the library other than the fused k_step path at the configs[1] size, for ncu captures
(tools/gpu_ncu_others.sh): k_reset, k_obs<int32> (reset / multi-frame), k_obs<int16>, k_ram, k_pack, k_order, k_flags."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from agarcl_b200 import OBS_I16, make_cfg  # noqa: E402
from agarcl_b200.batch import Batch  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(0)


def run(steps, **kw):
    w = dict(bench.WORKLOAD)
    w.update(kw)
    b = Batch(make_cfg(n_instances=N, **w))
    b.seed(np.arange(N, dtype=np.uint64) + np.uint64(1))
    b.reset()                                    # k_reset + k_obs
    for _ in range(steps):
        b.set_actions(rng.uniform(-1, 1, size=(N * b.A, 2)).astype(np.float32), rng.integers(0, 3, size=N * b.A).astype(np.int32))
        b.step()
    torch.cuda.synchronize()
    return b


b = run(3, num_frames=2)                         # k_step per tick + k_obs<int32> per frame
b.sync_mirror()                                  # k_pack<int32>
b.flags()                                        # k_flags
b.close()
b = run(3, num_frames=2, obs_dtype=OBS_I16)      # k_obs<int16>
b.sync_mirror()                                  # k_pack<int16>
b.close()
b = run(3, ram_obs=2, num_bots=8, num_viruses=10)  # k_ram (configs[2] roster), k_order
b.close()
print("done")
