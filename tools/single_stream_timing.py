#!/usr/bin/env python
"""Device-time one large raw-DEFLATE stream (BASELINE configs C1 / C5 shape) through the C ABI, in
single-stream mode (speculative warp decode, huff_stream.cuh) and in batch mode (one lane), and
check the bytes against zlib.  Usage: python tools/single_stream_timing.py [MiB ...]"""
import os
import sys
import time
import zlib

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import deflate_tools as T  # noqa: E402
import starflate_b200 as sfb  # noqa: E402


text = T.big_text


def run(mib, modes=("blocks", "warp", "lane"), reps=3):
    plain = text(mib << 20, 4242)
    t0 = time.time()
    comp = T.raw_deflate(plain, 6)
    t_comp = time.time() - t0
    dev = torch.device("cuda:0")
    src = torch.frombuffer(bytearray(comp + b"\0" * 64), dtype=torch.uint8).to(dev)
    dst = torch.empty(len(plain) + 256, dtype=torch.uint8, device=dev)
    one = lambda v, dt: torch.tensor([v], dtype=dt, device=dev)
    src_off, src_len = one(0, torch.int64), one(len(comp), torch.int64)
    dst_off, dst_cap = one(0, torch.int64), one(len(plain), torch.int64)
    status, written = one(9, torch.uint8), one(0, torch.int64)
    want = torch.frombuffer(bytearray(plain), dtype=torch.uint8).to(dev)
    for mode in modes:
        # blocks: finder + a warp per block + pointer jumping | warp: one warp front to back + pointer
        # jumping | lane: the batch kernels (one lane, one warp)
        os.environ["SFB200_STREAM_MODE"] = "0" if mode == "lane" else "1"
        os.environ["SFB200_BLOCKS"] = "1" if mode == "blocks" else "0"
        if mode == "lane" and mib > 64:
            continue
        ctx = sfb.Context(0)
        best = None
        for _ in range(reps):
            dst.fill_(0xA5)
            torch.cuda.synchronize()
            ctx.decompress_batch_device(src, src_off, src_len, dst, dst_off, dst_cap, status, written)
            torch.cuda.synchronize()
            p = ctx.last_pass_ms()
            if best is None or p[1] + p[2] < best[1] + best[2]:
                best = p
        ok = int(status.item()) == 0 and int(written.item()) == len(plain) and bool(torch.equal(dst[:len(plain)], want))
        ms = best[1] + best[2]
        print(f"{mib} MiB plain, {len(comp)} B deflate (zlib {t_comp:.1f} s): route={mode} "
              f"pass1 {best[1]:.2f} ms pass2 {best[2]:.2f} ms -> {len(plain) / ms / 1e6:.3f} GB/s, "
              f"bit-exact vs zlib input: {ok}", flush=True)
        ctx.close()
        if not ok:
            bad = torch.nonzero(dst[:len(plain)] != want).flatten()
            print("status", int(status.item()), "written", int(written.item()), "mismatches", bad.numel(),
                  "first", bad[:8].tolist(), "last", bad[-4:].tolist())
        assert ok
    assert zlib.decompress(comp, -15) == plain


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    routes = [a.split("=", 1)[1].split(",") for a in sys.argv[1:] if a.startswith("--routes=")]
    reps = [int(a.split("=", 1)[1]) for a in sys.argv[1:] if a.startswith("--reps=")]
    for m in [int(a) for a in args] or [1, 16]:
        run(m, *(routes[:1] or [("blocks", "warp", "lane")]), reps=(reps or [3])[0])
