#!/usr/bin/env python3
"""Fixture generator: a file -> raw DEFLATE (RFC 1951, no zlib/gzip wrapper) on stdout.

Takes the flags of the reference's fixture tool (/root/reference/tools/deflate_compress.py:7-19,
BASELINE.json configs[0]: `--src FILE [--fixed]`, default level, zlib's own block choices) so that
the same command lines produce the same bytes, and adds what the parity corpus needs:

  --level N     zlib level 0-9 (0: stored blocks)            --out FILE   write there instead of stdout
  --gpu         compress with this repository's GPU compressor (sfb200_compress) instead of zlib
  --check       decode the result again (zlib, and the GPU decoder when --gpu) and compare

  python tools/deflate_fixture.py --src src/test/starfleet.html > starfleet.html.dynamic
  python tools/deflate_fixture.py --src src/test/starfleet.html --fixed > starfleet.html.fixed"""
from __future__ import annotations

import argparse
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main() -> int:
    ap = argparse.ArgumentParser(prog="deflate_fixture", description=__doc__.splitlines()[0])
    ap.add_argument("--src", required=True, help="path to input file")
    ap.add_argument("--fixed", action="store_true", help="fixed-Huffman blocks only (zlib Z_FIXED)")
    ap.add_argument("--level", type=int, default=-1, help="zlib level (default: zlib's default, 6)")
    ap.add_argument("--gpu", action="store_true", help="use the GPU compressor of this repository")
    ap.add_argument("--check", action="store_true", help="decode again and compare")
    ap.add_argument("--out", default="", help="output file (default: stdout)")
    a = ap.parse_args()
    with open(a.src, "rb") as f:
        data = f.read()
    if a.gpu:
        import starflate_b200 as S
        ctx = S.Context(0)
        st, comp = ctx.compress(data)
        if st != 0:
            print(f"sfb200_compress: status {st}", file=sys.stderr)
            return 1
        if a.check:
            st2, back, wr = ctx.decompress(comp, len(data))
            assert (st2, wr) == (0, len(data)) and back == data, "GPU decode of the GPU compressor's output differs"
        ctx.close()
    else:
        from tests import deflate_tools as T
        comp = T.raw_deflate(data, a.level, zlib.Z_FIXED if a.fixed else zlib.Z_DEFAULT_STRATEGY)
    if a.check:
        assert zlib.decompress(comp, -15) == data, "zlib decode differs"
    out = open(a.out, "wb") if a.out else sys.stdout.buffer
    out.write(comp)
    if a.out:
        out.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
