#!/usr/bin/env python3
"""A/B timing of kernel variants in ONE process (one gpurun call): for every workload, generate the
batch once, then for every variant set its environment switches, create a context and time the
device-resident batch call (CUDA events inside the C ABI: clear | pass 1 | pass 2).  Every variant's
output is checked (status, sizes, CRC-32 of a sample of streams against the generator's).

  python tools/ab_bench.py --workloads c2,c3,c4 --variants "v2:;v1:SFB200_LZ_V1=1" [--steps 5]
      [--unique 4096] [--streams 0] [--out gpurun_out/ab.jsonl]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c2")
    ap.add_argument("--variants", default="base:")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--unique", type=int, default=4096)
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import starflate_b200 as S
    from starflate_b200 import build
    build.build_all()
    dev = torch.device("cuda", 0)
    variants = []
    for v in args.variants.split(";"):
        name, _, envs = v.partition(":")
        variants.append((name, dict(e.split("=", 1) for e in envs.split(",") if e)))
    all_keys = sorted({k for _, e in variants for k in e})
    out = open(args.out, "a") if args.out else None
    defaults = {"c1": 1, "c2": 65536, "c3": 1048576, "c4": 16384, "c5": 1, "html": 65536,
                "c3s": 349525, "c3f": 349525, "c3d": 349525}
    for wl in args.workloads.split(","):
        n_streams = args.streams or defaults[wl]
        uniq = {"c4": min(args.unique, 256), "c1": 1, "c5": 1}.get(wl, args.unique)
        w = bench.make_workload(wl, n_streams, min(uniq, n_streams), 0)
        n = w["n"]
        to_i64 = lambda a: torch.from_numpy(a.view(np.int64)).to(dev)
        d_src = torch.from_numpy(w["src"]).to(dev)
        d_so, d_sl, d_do, d_dc = to_i64(w["src_off"]), to_i64(w["src_len"]), to_i64(w["dst_off"]), to_i64(w["dst_cap"])
        d_dst = torch.zeros(w["total_out"] + 64, dtype=torch.uint8, device=dev)
        d_st = torch.zeros(n, dtype=torch.uint8, device=dev)
        d_wr = torch.zeros(n, dtype=torch.int64, device=dev)
        sample = np.unique(np.linspace(0, n - 1, 48).astype(np.int64))
        for name, env in variants:
            for k in all_keys:
                os.environ.pop(k, None)
            os.environ.update(env)
            ctx = S.Context(0)
            d_dst.fill_(0xA5)
            res = []
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tot = []
            for i in range(args.steps + 2):
                e0.record()
                ctx.decompress_batch_device(d_src, d_so, d_sl, d_dst, d_do, d_dc, d_st, d_wr)
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    res.append(ctx.last_pass_ms())
                    tot.append(e0.elapsed_time(e1))
            ok = int(d_st.max()) == 0 and bool(torch.equal(d_wr, d_dc))
            qdbg = None
            if env.get("SFB200_QUEUE", "0") != "0":  # the queue's counters: tail, head, final_tail, done, dbg[0..3]
                import ctypes
                buf = (ctypes.c_ulonglong * 8)()
                if ctx.lib.sfb200_debug_queue(ctx.h, buf) == 0:
                    qdbg = [int(x) for x in buf]
            for i in sample:
                o, c = int(w["dst_off"][i]), int(w["dst_cap"][i])
                ok = ok and zlib.crc32(d_dst[o:o + c].cpu().numpy().tobytes()) == int(w["crc"][i])
            if not res:
                ctx.close()
                continue
            m = np.median(np.array(res), axis=0)
            t = float(np.median(tot))
            line = {"workload": wl, "variant": name, "env": env, "ok": bool(ok), "n": n, "total_out": w["total_out"],
                    "total_in": w["total_in"], "clear_ms": float(m[0]), "pass1_ms": float(m[1]), "pass2_ms": float(m[2]),
                    "step_ms": t, "gbs": w["total_out"] / t / 1e6, "launch": ctx.launch_info(), "queue": qdbg}
            print(json.dumps(line), flush=True)
            if out:
                out.write(json.dumps(line) + "\n")
                out.flush()
            ctx.close()
        del d_src, d_dst
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
