"""Multi-GPU plumbing: faces / frames are independent, so they shard across ranks with no collective on the data
path.  torch.distributed is used for exactly two things: the barrier around a timed region and the
max-over-ranks of its duration; results are gathered on the host of rank 0."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of ceil(n / world) units for `rank` (the last ranks may get fewer, possibly none)."""
    per = -(-n // world) if world > 0 else n
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def shard_frames(image_of_box: np.ndarray, n_images: int, rank: int, world: int) -> np.ndarray:
    """Indices of the boxes whose frame belongs to `rank` (frames, not boxes, are the sharding unit: a frame is uploaded once)."""
    lo, hi = shard_range(n_images, rank, world)
    iob = np.asarray(image_of_box)
    return np.nonzero((iob >= lo) & (iob < hi))[0]


def max_over_ranks(value: float, dist=None, device=None) -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_records(local: np.ndarray, dist=None) -> np.ndarray | None:
    """Concatenates every rank's result records on rank 0 (host side; ~300 B per face)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    parts = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local, parts, dst=0)
    return np.concatenate(parts) if parts is not None else None
