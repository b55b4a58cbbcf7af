"""CPU, world_size 2 over gloo: the N>1 host logic — contiguous instance shards that cover the batch exactly
once, seeds keyed by global instance index (results independent of the GPU count), and the optional
gather of observation shards to a learner rank.  No collective is on the step path."""
import os
import socket
import subprocess
import sys

import numpy as np

from agarcl_b200.dist import shard_range, shard_seeds

WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["AGARCL_ROOT"])
from agarcl_b200.dist import shard_range, shard_seeds, gather_to_learner
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N = 1001
lo, hi = shard_range(N, world, rank)
seeds = shard_seeds(7, N, world, rank)
assert seeds.tolist() == list(range(7 + lo, 7 + hi))
# every rank contributes its shard bounds; together they must tile [0, N)
t = torch.tensor([lo, hi], dtype=torch.int64)
out = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
dist.all_gather(out, t)
bounds = sorted((int(o[0]), int(o[1])) for o in out)
assert bounds[0][0] == 0 and bounds[-1][1] == N and all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
# learner gather of equally sized obs shards (fake obs: value = global instance index)
per = 8
obs = (torch.arange(per, dtype=torch.int32) + rank * per).view(per, 1, 1, 1).expand(per, 2, 4, 4).contiguous()
g = gather_to_learner(obs, dst=0)
if rank == 0:
    assert g.shape == (per * world, 2, 4, 4) and g[:, 0, 0, 0].tolist() == list(range(per * world))
else:
    assert g is None
# max-over-ranks of a device time (what bench.py does)
ms = torch.tensor([10.0 + rank], dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
assert float(ms) == 10.0 + world - 1
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_properties():
    for n in (0, 1, 7, 4096, 65536, 1001):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    assert shard_seeds(100, 10, 2, 1).tolist() == [105, 106, 107, 108, 109]


def test_two_process_gloo(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, AGARCL_ROOT=root, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o
