#!/usr/bin/env python3
"""Token statistics of the synthetic corpora (design aid for the resolve pass): a plain-Python
inflate that only parses — literals per match, match length / distance histograms, how many
matched bytes have their source within N bytes.  Usage: token_stats.py [kind] [size] [streams]"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import deflate_tools as T  # noqa: E402


def build(lens):
    """code lengths -> dict (len, code) -> symbol (canonical)"""
    codes = T.canonical_codes(lens)
    return {(n, c): s for s, (c, n) in codes.items()}


class Bits:
    def __init__(self, data: bytes):
        self.v = int.from_bytes(data, "little")
        self.p = 0

    def get(self, n):
        r = (self.v >> self.p) & ((1 << n) - 1)
        self.p += n
        return r

    def sym(self, tab):
        c = 0
        for n in range(1, 16):
            c = (c << 1) | ((self.v >> self.p) & 1)
            self.p += 1
            s = tab.get((n, c))
            if s is not None:
                return s
        raise ValueError("bad code")


def tokens(data: bytes):
    """-> list of (pos, length, dist) for matches, total size, literal count, blocks"""
    b = Bits(data)
    pos = 0
    out = []
    lits = 0
    blocks = 0
    while True:
        final = b.get(1)
        typ = b.get(2)
        blocks += 1
        if typ == 0:
            b.p = (b.p + 7) & ~7
            n = b.get(16)
            b.get(16)
            b.p += 8 * n
            pos += n
            lits += n
        else:
            if typ == 1:
                lit = build(T.FIXED_LIT_LENS)
                dist = build(T.FIXED_DIST_LENS)
            else:
                hlit, hdist, hclen = b.get(5) + 257, b.get(5) + 1, b.get(4) + 4
                cl = [0] * 19
                for i in range(hclen):
                    cl[T.CL_ORDER[i]] = b.get(3)
                clt = build(cl)
                lens = []
                while len(lens) < hlit + hdist:
                    s = b.sym(clt)
                    if s < 16:
                        lens.append(s)
                    elif s == 16:
                        lens += [lens[-1]] * (3 + b.get(2))
                    elif s == 17:
                        lens += [0] * (3 + b.get(3))
                    else:
                        lens += [0] * (11 + b.get(7))
                lit = build(lens[:hlit])
                dist = build(lens[hlit:hlit + hdist])
            while True:
                s = b.sym(lit)
                if s < 256:
                    pos += 1
                    lits += 1
                elif s == 256:
                    break
                else:
                    i = s - 257
                    ln = T.LEN_BASE[i] + b.get(T.LEN_EXTRA[i])
                    d = b.sym(dist)
                    dd = T.DIST_BASE[d] + b.get(T.DIST_EXTRA[d])
                    out.append((pos, ln, dd))
                    pos += ln
        if final:
            break
    return out, pos, lits, blocks


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "dynamic"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    tot_b = tot_l = tot_m = tot_c = tot_blocks = 0
    lens, dists, overl = [], [], 0
    chunk_near = {128: [0, 0], 256: [0, 0], 512: [0, 0]}
    for i in range(n):
        plain, comp = T.make_stream(kind, size, 1_000_003 + i)
        tk, total, lits, blocks = tokens(comp)
        assert total == len(plain)
        tot_b += total
        tot_l += lits
        tot_m += len(tk)
        tot_c += len(comp)
        tot_blocks += blocks
        for (p, ln, d) in tk:
            lens.append(ln)
            dists.append(d)
            if d < ln:
                overl += 1
        for ch in chunk_near:
            # chunks with at least one match whose source reaches into the same chunk
            flag = np.zeros((total + ch - 1) // ch, dtype=bool)
            for (p, ln, d) in tk:
                for c in range(p // ch, (p + ln - 1) // ch + 1):
                    lo = max(p, c * ch)
                    hi = min(p + ln, (c + 1) * ch)
                    # sources of bytes [lo, hi): from lo - d (periodic); in-chunk iff lo-d+... >= c*ch
                    if hi - 1 - d >= c * ch:
                        flag[c] = True
            chunk_near[ch][0] += int(flag.sum())
            chunk_near[ch][1] += len(flag)
    lens = np.array(lens)
    dists = np.array(dists)
    print(f"{kind} x{n} size {size}: ratio {tot_b / tot_c:.2f}, blocks/stream {tot_blocks / n:.1f}")
    print(f"tokens/stream {(tot_l + tot_m) / n:.0f}  literals {tot_l / (tot_l + tot_m):.2%} of tokens, "
          f"{tot_l / tot_b:.2%} of bytes; bytes/token {tot_b / (tot_l + tot_m):.2f}; bits/token {8 * tot_c / (tot_l + tot_m):.2f}")
    print(f"match len mean {lens.mean():.2f} median {np.median(lens):.0f} p90 {np.percentile(lens, 90):.0f} max {lens.max()}; "
          f"len==3 {np.mean(lens == 3):.2%}; overlapping (d<len) {overl / len(lens):.3%}")
    for t in (4, 16, 64, 128, 256, 512, 1024, 4096, 16384):
        print(f"  dist < {t:6d}: {np.mean(dists < t):.2%} of matches, {lens[dists < t].sum() / tot_b:.2%} of bytes")
    for ch, (a, b2) in chunk_near.items():
        print(f"  chunks of {ch} B with an in-chunk source: {a / b2:.2%}")


if __name__ == "__main__":
    main()
