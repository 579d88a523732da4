"""Multi-GPU use of the batch decoder: streams are independent, so a batch is cut into contiguous
ranges of streams, one per GPU / rank, balanced by the bytes each stream moves (compressed read +
decompressed written — SURVEY.md §8e).  No data-path collective exists or is needed; the helpers
here only do the host-side arithmetic and the max-over-ranks reduction of a timing."""
from __future__ import annotations

import numpy as np


def partition_streams(src_len, dst_cap, world: int):
    """-> list of `world` (first, count) ranges, contiguous and covering all streams, with
    sum(src_len + dst_cap) per range as even as a contiguous cut allows.  The arithmetic is the
    library's own (sfb200_partition_streams, which sfb200_decompress_batch_host_multi uses to cut a
    batch across the GPUs of a box): host code, no device needed."""
    import ctypes as C

    from . import load_library
    src_len = np.ascontiguousarray(src_len, dtype=np.uint64)
    dst_cap = np.ascontiguousarray(dst_cap, dtype=np.uint64)
    n = len(src_len)
    assert len(dst_cap) == n and world >= 1
    lib = load_library()
    u64p = C.POINTER(C.c_uint64)
    lib.sfb200_partition_streams.argtypes = [u64p, u64p, C.c_uint64, C.c_int, u64p]
    lib.sfb200_partition_streams.restype = None
    cut = np.zeros(world + 1, dtype=np.uint64)
    lib.sfb200_partition_streams(src_len.ctypes.data_as(u64p), dst_cap.ctypes.data_as(u64p), n, world,
                                 cut.ctypes.data_as(u64p))
    return [(int(cut[r]), int(cut[r + 1] - cut[r])) for r in range(world)]


def max_over_ranks(value: float, device=None) -> float:
    """The slowest rank's value (timings of a sharded job are the max over ranks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
