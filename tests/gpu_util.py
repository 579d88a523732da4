"""Helpers for the `-m gpu` tests: move a tests.deflate_tools.Batch to the device and run it
through the C ABI (device-resident entry point)."""
from __future__ import annotations

import numpy as np
import torch

K_MUL = np.uint64(0x9E3779B97F4A7C15)


def run_device(ctx, b, fill=0xA5, want_written=True):
    dev = torch.device("cuda", ctx.device)
    src = torch.from_numpy(b.src).to(dev)
    dst = torch.full((b.dst_total,), fill, dtype=torch.uint8, device=dev)
    as_i64 = lambda a: torch.from_numpy(a.view(np.int64)).to(dev)
    src_off, src_len = as_i64(b.src_off), as_i64(b.src_len)
    dst_off, dst_cap = as_i64(b.dst_off), as_i64(b.dst_cap)
    status = torch.full((b.n,), 0xEE, dtype=torch.uint8, device=dev)
    written = torch.full((b.n,), -1, dtype=torch.int64, device=dev) if want_written else None
    ctx.decompress_batch_device(src, src_off, src_len, dst, dst_off, dst_cap, status, written)
    torch.cuda.synchronize(dev)
    return (status.cpu().numpy(), written.cpu().numpy().view(np.uint64) if want_written else None,
            dst.cpu().numpy())


def host_checksum(data: bytes) -> int:
    """sum_j (b_j + 1) * ((K*(j+1)) | 1) mod 2^64 — same as sfb200_checksum_batch_device."""
    a = np.frombuffer(data, dtype=np.uint8).astype(np.uint64)
    with np.errstate(over="ignore"):
        j = np.arange(1, len(a) + 1, dtype=np.uint64)
        w = (K_MUL * j) | np.uint64(1)
        return int(((a + np.uint64(1)) * w).sum(dtype=np.uint64))
