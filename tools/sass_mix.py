#!/usr/bin/env python3
"""Static instruction mix of one kernel from `cuobjdump -sass`: instructions per basic block (split
at labels / branches) classified by issue pipe (B300_MICROARCH.md: the ALU pipe — IADD3 / LOP3 /
SHF / PRMT / ISETP / SEL / IMNMX ... — and the FMA pipe — IMAD / FFMA ... — each take one warp
instruction per two cycles per scheduler).  Design aid: which loop is ALU-pipe bound, before any
GPU time is spent.   usage: sass_mix.py <lib.so> <kernel-name-substring> [--blocks]"""
from __future__ import annotations

import re
import subprocess
import sys
from collections import Counter

ALU = ("IADD3", "IADD", "LOP3", "LOP", "SHF", "SHL", "SHR", "PRMT", "ISETP", "SEL", "IMNMX", "VIMNMX", "LEA", "PLOP3",
       "FSETP", "FMNMX", "ICMP", "SGXT", "BMSK", "FSEL", "IABS", "VIADD", "VIADDMNMX", "MOV", "CS2R", "P2R", "R2P", "FCHK")
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2", "IDP")
XU = ("FLO", "POPC", "BREV", "MUFU", "I2F", "F2I", "I2I", "F2F", "I2FP", "F2IP")
LSU = ("LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOM", "ATOMS", "ATOMG", "RED", "LDGSTS", "LD", "ST", "LDSM", "CCTL",
       "MEMBAR", "ERRBAR", "LDGDEPBAR", "DEPBAR", "UBLKCP", "SYNCS", "FENCE", "UTMALDG", "UTMASTG", "REDUX", "LDC", "LDCU", "MATCH")
CBU = ("BRA", "BSSY", "BSYNC", "EXIT", "RET", "CALL", "BREAK", "WARPSYNC", "BAR", "YIELD", "NANOSLEEP", "JMP", "BRX", "JMX", "BMOV")
UNI = ("S2UR", "R2UR", "UMOV", "ULOP3", "UIADD3", "USHF", "ULEA", "UISETP", "UIMAD", "USEL", "UPRMT", "UFLO", "UPOPC", "VOTEU",
       "UPLOP3", "ULDC", "UBREV", "UP2UR", "UR2UP", "S2R", "VOTE", "SHFL")


def pipe(op: str) -> str:
    base = op.split(".")[0]
    for name, tab in (("alu", ALU), ("fma", FMA), ("xu", XU), ("lsu", LSU), ("cbu", CBU), ("uni", UNI)):
        if base in tab:
            return name
    if base.startswith("U"):
        return "uni"
    return "other:" + base


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    show_blocks = "--blocks" in sys.argv
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, funcs = None, {}
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", line)
        if m:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    for name, ins in funcs.items():
        if pat not in name:
            continue
        # branch targets -> block starts
        starts = {ins[0][0]}
        for k, (addr, text) in enumerate(ins):
            t = re.search(r"\b(BRA|BSSY|CALL\S*|BRX|JMP)\S*\s.*?(0x[0-9a-f]+)", text)
            if t:
                starts.add(int(t.group(2), 16))
            if re.search(r"\b(BRA|EXIT|RET|BRX|JMP|BSYNC)\b", text) and k + 1 < len(ins):
                starts.add(ins[k + 1][0])
        total = Counter()
        blocks = []
        cur_b = None
        for addr, text in ins:
            if addr in starts:
                cur_b = [addr, Counter(), 0, []]
                blocks.append(cur_b)
            t = text
            if t.startswith("@"):
                t = t.split(None, 1)[1]
            op = t.split()[0]
            if op == "NOP":
                continue
            p = pipe(op)
            total[p] += 1
            cur_b[1][p] += 1
            cur_b[2] += 1
            cur_b[3].append(op.split(".")[0])
        print(f"== {name}: {sum(total.values())} instructions  " + "  ".join(f"{k}={v}" for k, v in total.most_common()))
        if show_blocks:
            for addr, c, n, ops in blocks:
                if n:
                    print(f"   {addr:#06x} n={n:4d}  " + " ".join(f"{k}={v}" for k, v in c.most_common()))


if __name__ == "__main__":
    main()
