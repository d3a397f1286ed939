"""Multi-GPU plumbing of the rollout: episodes are independent, so every rank owns its own
episodes, weight replica and random stream and there is NO data-path collective (the reference
already runs its environments in separate processes, l_dgn.py:137).  torch.distributed is only
used to agree on the timing (max over ranks) and to add up the work (sum over ranks)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_tuples(arrays, rank: int, episodes_per_rank: int):
    """Every rank reads the same reset-tuple pool but starts at a different offset, so ranks do
    not replay each other's episodes.  ``arrays``: tuple of numpy arrays with the pool on axis 0."""
    shift = rank * (episodes_per_rank // 2)
    return tuple(np.roll(a, -shift, axis=0) for a in arrays)


def reduce_job(ms: float, units: float, device=None):
    """-> (max over ranks of ms, sum over ranks of units).  Works on gloo (CPU) and nccl."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms), float(units)
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


def whole_job_throughput(ms: float, units: float, device=None) -> float:
    """units per second of the whole job: all ranks' units / the slowest rank's time."""
    ms_max, total = reduce_job(ms, units, device)
    return total / (ms_max / 1e3) if ms_max > 0 else 0.0
