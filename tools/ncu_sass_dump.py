#!/usr/bin/env python3
"""Dump the SASS of one kernel annotated with ncu per-instruction counts and source lines.
usage: ncu_sass_dump.py sass.csv dis.txt kernel_substring [min_exec]"""
import csv, re, sys
sass_csv, dis_txt, kname = sys.argv[1:4]
min_exec = int(sys.argv[4]) if len(sys.argv) > 4 else 0
line_of = {}; cur = None; infn = False; labels = {}
pending_label = None
for l in open(dis_txt):
    if l.startswith("//---") and ".text." in l:
        infn = kname in l
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = int(m.group(2)); continue
    m = re.match(r"^(\.L_x_\d+):", l)
    if m: pending_label = m.group(1); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off = int(m.group(1), 16)
        line_of[off] = (cur, m.group(2).strip())
        if pending_label: labels[off] = pending_label; pending_label = None
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
base = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    addr = int(r[ix["Address"]], 16)
    if base is None: base = addr
    off = addr - base
    inst = int(r[ix["Instructions Executed"]] or 0)
    if inst < min_exec: continue
    thr = float(r[ix["Avg. Threads Executed"]] or 0)
    samp = int(r[ix["# Samples"]] or 0)
    ln, txt = line_of.get(off, (0, r[ix["Source"]]))
    lab = labels.get(off, "")
    print(f"{off:05x} {lab:9s} x{inst:9d} t{thr:4.1f} s{samp:6d} L{ln:4d} | {txt}")
