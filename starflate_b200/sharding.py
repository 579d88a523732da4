"""Multi-GPU use of the batch decoder: streams are independent, so a batch is cut into contiguous
ranges of streams, one per GPU / rank, balanced by the bytes each stream moves (compressed read +
decompressed written — SURVEY.md §8e).  No data-path collective exists or is needed; the helpers
here only do the host-side arithmetic and the max-over-ranks reduction of a timing."""
from __future__ import annotations

import numpy as np


def partition_streams(src_len, dst_cap, world: int):
    """-> list of `world` (first, count) ranges, contiguous and covering all streams, with
    sum(src_len + dst_cap) per range as even as a contiguous cut allows."""
    src_len = np.asarray(src_len, dtype=np.uint64)
    dst_cap = np.asarray(dst_cap, dtype=np.uint64)
    n = len(src_len)
    assert len(dst_cap) == n and world >= 1
    cost = np.cumsum((src_len + dst_cap).astype(np.float64))
    total = float(cost[-1]) if n else 0.0
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        cut = int(np.searchsorted(cost, target, side="left")) if n else 0
        if cut < n and n:  # put the boundary on the nearer side of the stream that straddles it
            before = float(cost[cut - 1]) if cut else 0.0
            if abs(float(cost[cut]) - target) <= abs(before - target):
                cut += 1
        bounds.append(max(bounds[-1], min(cut, n)))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1] - bounds[r]) for r in range(world)]


def max_over_ranks(value: float, device=None) -> float:
    """The slowest rank's value (timings of a sharded job are the max over ranks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
