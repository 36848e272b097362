"""Multi-GPU sharding of the MSM: one process per GPU, contiguous point ranges, one exchange.

The reference is single-process; this is the new K7 step.  sum_i s_i P_i over disjoint
index ranges is independent per rank; the only exchange is the "all-reduce" of the per-rank partial G1
accumulators.  NCCL has no curve-addition reduction operator, so the all-reduce is realised as an
all-gather of the 144-byte normalised Jacobian partials followed by world_size - 1 additions done
identically on every rank (``combine`` = Context.g1_sum on the device).
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, end) of the contiguous point range owned by ``rank`` (even split, remainder to low ranks)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_g1(partial: np.ndarray, combine: Callable[[np.ndarray], np.ndarray], group=None, device=None) -> np.ndarray:
    """partial: (18,) uint64 Jacobian point of this rank -> sum over all ranks (same on every rank).

    ``device``: torch device the collective runs on ("cuda:k" for NCCL, None/"cpu" for gloo)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return partial
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(partial).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    allp = torch.stack(out).cpu().numpy().view(np.uint64)
    return combine(allp)


def sharded_msm(ctx, srs_shard, scalars_shard, group=None, device=None) -> np.ndarray:
    """Each rank: local MSM over its shard (device), then the exchange."""
    partial = ctx.msm(srs_shard, scalars_shard)
    return allreduce_g1(partial, ctx.g1_sum, group=group, device=device)
