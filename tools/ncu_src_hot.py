#!/usr/bin/env python3
"""Rank CUDA source lines of an ncu report (--import-source on, -lineinfo) by stall samples.

  python tools/ncu_src_hot.py prof.ncu-rep [top] [kernel-substring]
Prints, per source line: samples, share, executed warp instructions, avg threads, top stalls.
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    kfilter = sys.argv[3] if len(sys.argv) > 3 else ""
    sort_key = sys.argv[4] if len(sys.argv) > 4 else "samples"   # or "inst"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    cur_file = cur_fn = ""
    agg = {}
    use = True
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            cur_fn = r[1]
            use = kfilter in cur_fn
            continue
        if r[0] == "Line No":
            hdr = r
            ix = {h: i for i, h in enumerate(hdr)}
            stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or not use or not r[0]:
            continue  # SASS rows have an empty line number
        key = (cur_file, int(r[0]))
        try:
            samples = int(r[ix["# Samples"]])
            inst = int(r[ix["Instructions Executed"]])
            thr = int(r[ix["Thread Instructions Executed"]])
        except (ValueError, KeyError):
            continue
        a = agg.setdefault(key, {"src": r[1].strip(), "samples": 0, "inst": 0, "thr": 0,
                                 "stalls": defaultdict(int)})
        a["samples"] += samples
        a["inst"] += inst
        a["thr"] += thr
        for h, i in stall_cols:
            try:
                a["stalls"][h] += int(r[i])
            except ValueError:
                pass
    tot = sum(a["samples"] for a in agg.values()) or 1
    tot_inst = sum(a["inst"] for a in agg.values()) or 1
    print(f"total samples {tot}, warp instructions {tot_inst}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][sort_key])[:top]:
        st = sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3]
        sts = " ".join(f"{h[6:]}={v}" for h, v in st if v)
        thr = a["thr"] / a["inst"] if a["inst"] else 0
        print(f"{key[0]}:{key[1]:<5d} {a['samples']:7d} {100*a['samples']/tot:5.1f}%  inst {a['inst']:>11d} "
              f"({100*a['inst']/tot_inst:4.1f}%) thr {thr:4.1f}  {sts:42s} | {a['src'][:90]}")


if __name__ == "__main__":
    main()
