#!/usr/bin/env python
"""compute-sanitizer driver for k_step at steady state.  The sanitizer slows kernels down 10-100x, so the
game age is produced natively first and carried over as state blobs:
    python tools/sanitize_run.py --prepare /tmp/states.npy [--instances 600] [--age 1500]
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_run.py --run /tmp/states.npy [--steps 3]
The run uploads the blobs into a fresh batch (configs[1] workload, Philox) and steps it through agarcl_batch_step_mirror, i.e.
the fused k_step with the host-mirror lists, k_order, and (first call) k_pack."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prepare")
    ap.add_argument("--run")
    ap.add_argument("--instances", type=int, default=600)
    ap.add_argument("--age", type=int, default=1500)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import torch
    import bench
    from agarcl_b200 import make_cfg
    from agarcl_b200._abi import StateView
    from agarcl_b200.batch import Batch
    if a.prepare:
        N = a.instances
        b = Batch(make_cfg(n_instances=N, **bench.WORKLOAD))
        b.seed(np.arange(N, dtype=np.uint64) + np.uint64(1))
        b.reset()
        rng = np.random.default_rng(0)
        for _ in range(a.age):
            b.set_actions(rng.uniform(-1, 1, size=(N, 2)).astype(np.float32), rng.integers(0, 3, size=N).astype(np.int32))
            b.step()
        torch.cuda.synchronize()
        blobs = np.stack([b.download_state(i).blob for i in range(N)])
        np.save(a.prepare, blobs)
        print("prepared", blobs.shape, "multi-cell players per instance", float(np.mean([(b.download_state(i).players["n_cells"] >= 2).sum() for i in range(0, N, 37)])))
        return
    blobs = np.load(a.run)
    N = blobs.shape[0]
    b = Batch(make_cfg(n_instances=N, **bench.WORKLOAD))
    b.seed(np.arange(N, dtype=np.uint64) + np.uint64(1))
    b.reset()
    for i in range(N):
        b.upload_state(i, StateView(b.layout, blobs[i].copy()))
    rng = np.random.default_rng(1)
    rew, done = np.zeros(N, np.float64), np.zeros(N, np.uint8)
    for _ in range(a.steps):
        m = b.step_mirror(rng.uniform(-1, 1, size=(N, 2)).astype(np.float32), rng.integers(0, 3, size=N).astype(np.int32), rew, done)
    torch.cuda.synchronize()
    ok = bool(np.array_equal(m, b.obs_tensor().cpu().numpy()))
    print("ran", a.steps, "steps of", N, "instances at tick", int(b.download_state(0).hdr["tick"]), "mirror == device:", ok, "flags", b.flags())
    b.close()


if __name__ == "__main__":
    main()
