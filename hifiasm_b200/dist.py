"""Multi-GPU plumbing: query reads shard across ranks, index + reads replicated,
no data-path collective.  torch.distributed is used only to
agree on the timing (max over ranks) and the totals."""
from __future__ import annotations


def shard_range(n_reads: int, rank: int, world: int):
    """Contiguous shard of read ids for `rank` (the last ranks may be short)."""
    per = (n_reads + world - 1) // world
    return min(n_reads, rank * per), min(n_reads, (rank + 1) * per)


def reduce_time_and_units(dev_ms: float, units: float, device=None):
    """-> (max over ranks of dev_ms, sum over ranks of units)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dev_ms, units
    t = torch.tensor([dev_ms], dtype=torch.float64, device=device)
    u = torch.tensor([units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])


def merge_shards(parts):
    """Concatenate per-rank (records, offsets) results (each offsets array is
    relative to its shard) into one (records, offsets) pair in read-id order."""
    import numpy as np
    recs = [p[0] for p in parts]
    offs = [np.zeros(1, np.uint64)]
    base = 0
    for r, o in parts:
        offs.append(o[1:].astype(np.uint64) + np.uint64(base))
        base += int(o[-1])
    return np.concatenate(recs), np.concatenate(offs)
