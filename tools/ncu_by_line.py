#!/usr/bin/env python3
"""Join an ncu SASS-level source page (csv) with nvdisasm line info to rank CUDA source lines
by executed warp-instructions and stall samples.

  ncu -i prof.ncu-rep --page source --csv > sass.csv
  cuobjdump -xelf all lib.so ; nvdisasm -g -c capi.sm_100a.cubin > dis.txt
  python tools/ncu_by_line.py sass.csv dis.txt kernel_name_substring [top]
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    sass_csv, dis_txt, kname = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    # nvdisasm: offset -> (file line)
    line_of = {}
    cur = None
    infn = False
    for l in open(dis_txt):
        if l.startswith("//---") and ".text." in l:
            infn = kname in l
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
    rows = list(csv.reader(open(sass_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ix = {h: i for i, h in enumerate(hdr)}
    base = None
    agg = defaultdict(lambda: [0, 0, 0, 0.0])  # instr, samples, thread-instr
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    stall_by_line = defaultdict(lambda: defaultdict(int))
    tot_i = tot_s = tot_t = 0
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[ix["Address"]], 16)
        if base is None:
            base = addr
        off = addr - base
        src, _ = line_of.get(off, (None, None))
        inst = int(r[ix["Instructions Executed"]] or 0)
        samp = int(r[ix["# Samples"]] or 0)
        thr = int(r[ix["Thread Instructions Executed"]] or 0)
        a = agg[src]
        a[0] += inst; a[1] += samp; a[2] += thr
        tot_i += inst; tot_s += samp; tot_t += thr
        for c in stall_cols:
            v = int(r[ix[c]] or 0)
            if v:
                stall_by_line[src][c] += v
    print(f"total warp-instr {tot_i}  samples {tot_s}  avg active threads {tot_t / max(tot_i, 1):.1f}")
    srcs = {}
    for (key, a) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if key is None:
            text = "?"
        else:
            f, ln = key
            if f not in srcs:
                try:
                    srcs[f] = open("starflate_b200/csrc/" + f).read().split("\n")
                except OSError:
                    srcs[f] = []
            text = srcs[f][ln - 1].strip() if ln - 1 < len(srcs[f]) else ""
        st = sorted(stall_by_line[key].items(), key=lambda kv: -kv[1])[:2]
        sts = " ".join(f"{k[6:]}={v}" for k, v in st)
        print(f"{(key[1] if key else 0):5d} inst={a[0] / tot_i * 100:5.1f}% samp={a[1] / max(tot_s,1) * 100:5.1f}% "
              f"thr={a[2] / max(a[0], 1):4.1f} [{sts}] | {text[:90]}")


if __name__ == "__main__":
    main()
