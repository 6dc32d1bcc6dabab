"""Multi-GPU sharding of independent NMPC instances.

Every NMPC instance is independent, so the batch is cut into contiguous index ranges, one per
rank / GPU, with NO collective on the solve path.  The only exchange is the final gather of the
small per-problem results (u0, status) to rank 0, done on the host side of torch.distributed
(gloo on CPU boxes, nccl on GPU boxes): it is outside every timed solve.
"""
from __future__ import annotations

import numpy as np


def shard_range(total: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of problems owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bucket_by_object(object_ids, world_size: int):
    """Sort problems by object id (block-uniform spline tables per warp), then cut into contiguous shards.

    Returns (perm, ranges): `perm` is the stable permutation to apply to every per-problem array,
    ranges[r] the [lo, hi) slice of the permuted batch owned by rank r.
    """
    ids = np.asarray(object_ids)
    perm = np.argsort(ids, kind="stable")
    return perm, [shard_range(len(ids), world_size, r) for r in range(world_size)]


def gather_to_rank0(local: np.ndarray, total: int, world_size: int, rank: int):
    """Final host gather: every rank contributes its [lo,hi) rows; rank 0 receives the full array."""
    import torch.distributed as dist

    if world_size == 1 or not dist.is_initialized():
        return local
    parts = [None] * world_size
    dist.all_gather_object(parts, np.ascontiguousarray(local))     # shard sizes may differ by one
    if rank != 0:
        return None
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == total
    return out
