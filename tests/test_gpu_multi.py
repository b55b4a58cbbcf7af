"""-m gpu, needs 2 GPUs (skipped otherwise): one process per GPU over NCCL -- the sharded batch reproduces the single-GPU
batch bit for bit, and dist.gather_to_learner moves REAL observation shards to the learner rank over NVLink."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["AGARCL_ROOT"])
from agarcl_b200.dist import shard_range, shard_seeds, gather_to_learner
from agarcl_b200.env import BatchedGridEnvironment
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
N = 64
lo, hi = shard_range(N, world, rank)
kw = dict(num_bots=6, arena_size=300, num_pellets=200, num_viruses=4)
env = BatchedGridEnvironment(hi - lo, device=rank, instance_base=lo, **kw)
env.seed(shard_seeds(11, N, world, rank))
env.reset()
rng = np.random.default_rng(3)
for st in range(25):
    dxdy = rng.uniform(-1, 1, size=(N, 2)).astype(np.float32)   # the same global action stream on every rank
    act = rng.integers(0, 3, size=N).astype(np.int32)
    obs, rew, done = env.step(torch.from_numpy(dxdy[lo:hi]).cuda(), torch.from_numpy(act[lo:hi]).cuda())
g_obs = gather_to_learner(obs.contiguous(), dst=0)
g_rew = gather_to_learner(rew.contiguous(), dst=0)
if rank == 0:
    assert g_obs.shape == (N, 8, 128, 128) and g_obs.is_cuda
    ref = BatchedGridEnvironment(N, device=0, instance_base=0, **kw)   # the whole batch on one GPU
    ref.seed(shard_seeds(11, N, 1, 0))
    ref.reset()
    rng = np.random.default_rng(3)
    for st in range(25):
        dxdy = rng.uniform(-1, 1, size=(N, 2)).astype(np.float32)
        act = rng.integers(0, 3, size=N).astype(np.int32)
        o1, r1, d1 = ref.step(torch.from_numpy(dxdy).cuda(), torch.from_numpy(act).cuda())
    assert torch.equal(g_obs, o1), "sharded observations differ from the single-GPU batch"
    assert torch.equal(g_rew, r1)
    ref.close()
else:
    assert g_obs is None
dist.barrier()
env.close()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_gpu_shards_equal_single_gpu_and_gather_over_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, AGARCL_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(w)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout
