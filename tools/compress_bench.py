#!/usr/bin/env python3
"""Compression on the C2 plaintexts: the GPU decoder's own output is the input.
Decode the C2 batch (65 536 x 64 KiB, checked against the generator's checksums), compress every
stream with sfb200_compress_batch_device (CUDA events), decode what it wrote with
sfb200_decompress_batch_device and compare the device checksums of that round trip with the first
decode's.  Prints one JSON line: compressed GB/s of INPUT, ratio, zlib level-6 ratio beside it.

  python tools/compress_bench.py [--streams 65536] [--unique 2048] [--steps 5]"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=65536)
    ap.add_argument("--unique", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--workload", default="c2")
    args = ap.parse_args()
    import torch
    import starflate_b200 as S
    from starflate_b200 import build
    build.build_all()
    dev = torch.device("cuda", 0)
    w = bench.make_workload(args.workload, args.streams, min(args.unique, args.streams), 0)
    n = w["n"]
    to_i64 = lambda a: torch.from_numpy(a.view(np.int64)).to(dev)
    ctx = S.Context(0)
    d_src = torch.from_numpy(w["src"]).to(dev)
    d_so, d_sl, d_do, d_dc = to_i64(w["src_off"]), to_i64(w["src_len"]), to_i64(w["dst_off"]), to_i64(w["dst_cap"])
    d_plain = torch.zeros(w["total_out"] + 64, dtype=torch.uint8, device=dev)
    d_st = torch.zeros(n, dtype=torch.uint8, device=dev)
    d_wr = torch.zeros(n, dtype=torch.int64, device=dev)
    ctx.decompress_batch_device(d_src, d_so, d_sl, d_plain, d_do, d_dc, d_st, d_wr)
    sums0 = torch.zeros(n, dtype=torch.int64, device=dev)
    ctx.checksum_batch_device(d_plain, d_do, d_wr, sums0)
    torch.cuda.synchronize()
    assert int(d_st.max()) == 0 and bool(torch.equal(d_wr, d_dc))
    assert np.array_equal(sums0.cpu().numpy().view(np.uint64), w["sum"]), "first decode is wrong"
    # compress: every stream gets the stored bound as room
    caps = w["dst_cap"] + np.uint64(5) * ((w["dst_cap"] + np.uint64(65534)) // np.uint64(65535)) + np.uint64(8)
    c_off = np.zeros(n, np.uint64)
    c_off[1:] = np.cumsum(caps)[:-1]
    d_comp = torch.zeros(int(caps.sum()) + 64, dtype=torch.uint8, device=dev)
    d_co, d_cc = to_i64(c_off), to_i64(caps)
    c_st = torch.zeros(n, dtype=torch.uint8, device=dev)
    c_wr = torch.zeros(n, dtype=torch.int64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for i in range(args.steps + 2):
        e0.record()
        ctx.compress_batch_device(d_plain, d_do, d_wr, d_comp, d_co, d_cc, c_st, c_wr)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(e0.elapsed_time(e1))
    assert int(c_st.max()) == 0
    comp_bytes = int(c_wr.sum())
    # decode what the compressor wrote, compare with the first decode
    d_back = torch.zeros(w["total_out"] + 64, dtype=torch.uint8, device=dev)
    b_st = torch.zeros(n, dtype=torch.uint8, device=dev)
    b_wr = torch.zeros(n, dtype=torch.int64, device=dev)
    ctx.decompress_batch_device(d_comp, d_co, c_wr, d_back, d_do, d_dc, b_st, b_wr)
    sums1 = torch.zeros(n, dtype=torch.int64, device=dev)
    ctx.checksum_batch_device(d_back, d_do, b_wr, sums1)
    torch.cuda.synchronize()
    ok = int(b_st.max()) == 0 and bool(torch.equal(b_wr, d_dc)) and bool(torch.equal(sums0, sums1))
    p1 = ctx.last_pass_ms()
    t = float(np.median(ms))
    print(json.dumps({"metric": "compression GB/s of input (device-timed), hashed LZ77 + one " +
                                ("fixed-Huffman" if os.environ.get("SFB200_COMPRESS_FIXED") == "1" else "dynamic- or fixed-Huffman") +
                                " block per stream",
                      "workload": w["desc"], "ok": ok, "ms": t, "value": w["total_out"] / t / 1e6, "unit": "GB/s",
                      "ratio": w["total_out"] / comp_bytes, "zlib6_ratio": w["total_out"] / w["total_in"],
                      "decode_of_own_output_ms": p1, "launch": ctx.launch_info()["kernel_launches"]}))
    ctx.close()


if __name__ == "__main__":
    main()
