#!/usr/bin/env python3
"""Static cost of the pass-1 token loop (huff_lanes_kernel) from the built library: instructions per
iteration by source line and issue pipe, assuming the loop body is straight-line predicated code
(it is: ncu's per-instruction execution counts agree, profiles/r02_*).  Lets a change to the loop be
judged without GPU time.

  python tools/loop_cost.py [lib.so] [--kernel SUBSTR] [--lines]
The loop is found as the address range between the first instruction attributed to the loop
condition line and the last backward branch to it; the rare paths inside it (exact slow path, long
codes) are excluded by source line.
NOTE (end of round 2): the rare-path line ranges below (deflate_lane.cuh 560-590, the markers in
huff_lanes.cuh) have not followed the last source changes: the total it prints includes the inlined
exact path.  The per-line execution counts of an ncu capture (tools/ncu_src_hot.py ... inst) are
what DESIGN.md's per-iteration figures come from.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from sass_mix import pipe  # noqa: E402


def disasm(lib: str) -> str:
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, check=True, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        return subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    lib = args[0] if args else os.path.join(ROOT, "starflate_b200", "_build", "libstarflate_b200.so")
    kern = "huff_lanes_kernelINS_3CfgILi8ELi6ELi96ELi4ELi2EEELb0"
    if "--kernel" in sys.argv:
        kern = sys.argv[sys.argv.index("--kernel") + 1]
    txt = disasm(lib)
    src_lines = open(os.path.join(ROOT, "starflate_b200", "csrc", "huff_lanes.cuh")).read().splitlines()
    loop_line = next(i + 1 for i, l in enumerate(src_lines) if "while (__any_sync(FULL, state == S_DECODE))" in l)
    rare_lo = next(i + 1 for i, l in enumerate(src_lines) if "// (a) a valid code longer than the tables hold" in l) - 1
    rare_hi = next(i + 1 for i, l in enumerate(src_lines) if "if (t.status == ST_SUCCESS) br.skip(used);" in l) - 1
    ins = []
    cur, infn = None, False
    for l in txt.splitlines():
        if l.startswith("//---") and ".text." in l:
            infn = kern in l
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
    if not ins:
        raise SystemExit("kernel not found")
    # the token loop: the address range spanned by the instructions attributed to the loop's own
    # source lines (the callees are inlined: their instructions lie in between)
    body_lo = loop_line + 1   # (the condition's own line also has instructions in front of the loop)
    body_hi = next(i + 1 for i, l in enumerate(src_lines) if "p_skip = mt ? value - 3u : 0u;" in l)
    own = [a for a, t, s in ins if s and s[0] == "huff_lanes.cuh" and body_lo <= s[1] <= body_hi and not (rare_lo <= s[1] <= rare_hi)]
    start, end = min(own), max(own)
    per_line = defaultdict(lambda: defaultdict(int))
    tot = defaultdict(int)
    rare = 0
    for a, t, s in ins:
        if a < start or a > end:
            continue
        if s and s[0] == "huff_lanes.cuh" and rare_lo <= s[1] <= rare_hi:
            rare += 1
            continue
        if s and s[0] == "deflate_lane.cuh" and 560 <= s[1] <= 590:  # long_decode (rare path only)
            rare += 1
            continue
        tt = t.split(None, 1)[1] if t.startswith("@") else t
        op = tt.split()[0]
        if op == "NOP":
            continue
        p = pipe(op)
        per_line[s][p] += 1
        tot[p] += 1
    n = sum(tot.values())
    print(f"token loop [{start:#x}, {end:#x}]: {n} instructions/iteration (+{rare} on rare paths)  " +
          "  ".join(f"{k}={v}" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])))
    if "--lines" in sys.argv:
        for s, pp in sorted(per_line.items(), key=lambda kv: -sum(kv[1].values())):
            print(f"  {str(s):36s} {sum(pp.values()):4d}  " + " ".join(f"{k}={v}" for k, v in sorted(pp.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main()
