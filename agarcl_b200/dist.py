"""Multi-GPU plumbing: one process per GPU, instances sharded by contiguous ranges, NO collective on
the step path (instances never interact: one Engine per environment, BaseEnvironment.hpp:346).
torch.distributed is used only for (a) the bench barrier / max-over-ranks and (b) the OPTIONAL gather
of observation shards to a learner rank, which is off the timed path.
"""
import numpy as np


def shard_range(n_total, world_size, rank):
    """Contiguous shard [lo, hi) of instance indices for `rank`; sizes differ by at most one."""
    assert 0 <= rank < world_size and n_total >= 0
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_seeds(base_seed, n_total, world_size, rank):
    """Seeds are keyed by GLOBAL instance index so results do not depend on the GPU count."""
    lo, hi = shard_range(n_total, world_size, rank)
    return np.arange(lo, hi, dtype=np.uint64) + np.uint64(base_seed)


def gather_to_learner(tensor, dst=0, group=None):
    """Gathers equally sized shards (obs / rewards / dones) on rank `dst` (NCCL on GPUs, gloo on CPU).
    Returns the concatenated tensor on dst, None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return tensor
    out = [torch.empty_like(tensor) for _ in range(world)] if rank == dst else None
    dist.gather(tensor, out, dst=dst, group=group)
    return torch.cat(out, dim=0) if rank == dst else None
