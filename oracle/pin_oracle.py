#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY — pins oracle/inflate_oracle.c to the UNMODIFIED reference and
writes the golden fixtures under tests/golden/.

Needs /root/reference (this container only).  For every case it runs
  oracle/_ref/classify_ndebug   reference, -O2 -DNDEBUG           -> status + dst fingerprint
  oracle/_ref/classify_san      reference, NDEBUG + ASan + UBSan   -> defined (R) / undefined (X)
  oracle/_ref/classify_assert   reference, asserts on              -> assert-only class (A)
and the C restatement, and requires
  san == R  =>  oracle.ref_undefined == 0, oracle status == reference status,
                oracle dst (whole capacity, 0xA5 pre-fill) fingerprint == reference's
  san == X  <=> oracle.ref_undefined == 1
Usage: python oracle/pin_oracle.py [--write]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import struct
import subprocess
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import bindings  # noqa: E402
from tests import deflate_tools as T  # noqa: E402
from tests import known_answers as K  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_HTML = "/root/reference/src/test/starfleet.html"


def build_bases():
    html = open(REF_HTML, "rb").read()
    assert hashlib.md5(html).hexdigest() == K.STARFLEET_MD5 and len(html) == K.STARFLEET_LEN
    bases = {
        # the reference's own fixtures: tools/deflate_compress.py call shape (default level)
        "starfleet_dynamic": (T.raw_deflate(html, -1), html),
        "starfleet_fixed": (T.raw_deflate(html, -1, zlib.Z_FIXED), html),
    }
    for kind, n, seed in [("dynamic", 4096, 11), ("fixed", 4096, 12), ("stored", 4096, 13),
                          ("multiblock", 12000, 14), ("repetitive", 70000, 15),
                          ("dynamic", 65536, 16)]:
        plain, comp = T.make_stream(kind, n, seed)
        bases[f"{kind}_{n}"] = (comp, plain)
    return bases


def crafted_cases(rng: np.random.Generator, count: int):
    """Semi-valid dynamic headers (random code-length sets, repeat codes at the edges) followed by
    random payload bits: exercises incomplete / over-subscribed codes and every header error."""
    out = []
    for k in range(count):
        w = T.BitWriter()
        hlit = int(rng.choice([0, 0, 0, 1, 3, 29, 30, 31]))
        hdist = int(rng.choice([0, 0, 1, 5, 29, 30, 31]))
        n_lit, n_dist = 257 + hlit, 1 + hdist
        mode = k % 4
        if mode == 0:      # sparse small alphabet, valid-ish
            lit = [0] * n_lit
            for s in rng.choice(n_lit, size=int(rng.integers(1, 12)), replace=False):
                lit[int(s)] = int(rng.integers(1, 6))
            lit[256] = int(rng.integers(0, 4))
            dist = [int(rng.integers(0, 4)) for _ in range(n_dist)]
        elif mode == 1:    # dense, probably over-subscribed or incomplete
            lit = [int(rng.integers(0, 16)) if rng.random() < 0.1 else 0 for _ in range(n_lit)]
            dist = [int(rng.integers(0, 16)) if rng.random() < 0.3 else 0 for _ in range(n_dist)]
        elif mode == 2:    # complete code from a real compressor-like profile
            lit = [0] * n_lit
            for s in range(n_lit):
                lit[s] = 9 if s < 256 else 7
            lit[:8] = [4] * 8
            dist = [5] * n_dist
        else:
            lit = [int(x) for x in rng.integers(0, 16, n_lit)]
            dist = [int(x) for x in rng.integers(0, 16, n_dist)]
        # express the two runs with occasional repeat symbols (legal and illegal placements)
        def run_syms(lens):
            syms, i = [], 0
            while i < len(lens):
                r = rng.random()
                if r < 0.08:
                    syms.append((16, int(rng.integers(0, 4)))); i += 3
                elif r < 0.16:
                    syms.append((17, int(rng.integers(0, 8)))); i += 3
                elif r < 0.22:
                    syms.append((18, int(rng.integers(0, 128)))); i += 11
                else:
                    syms.append((lens[i], 0)); i += 1
            return syms
        if rng.random() < 0.5:
            cl_symbols = [run_syms(lit), run_syms(dist)]
        else:
            cl_symbols = None
        hclen = int(rng.choice([19, 19, 19, 18, 12, 4]))
        cl_lens = [5] * 19 if rng.random() < 0.7 else [int(x) for x in rng.integers(0, 8, 19)]
        try:
            T.dynamic_header(w, bool(rng.integers(0, 2)), lit, dist, cl_lens, cl_symbols, hclen)
        except KeyError:
            # a needed CL symbol has no code under this cl_lens: fall back to the flat CL code
            w = T.BitWriter()
            T.dynamic_header(w, bool(rng.integers(0, 2)), lit, dist, [5] * 19, cl_symbols, 19)
        payload = rng.integers(0, 256, int(rng.integers(0, 24)), dtype=np.uint8).tobytes()
        if rng.random() < 0.3:
            payload = bytes(len(payload))
        stream = w.tobytes() + payload
        if rng.random() < 0.15:
            stream = stream[: int(rng.integers(1, len(stream) + 1))]
        out.append((stream, int(rng.choice([0, 4, 64, 600]))))
    return out


def build_families(bases):
    rng = np.random.default_rng(20240607)
    fams = []
    fams.append({"name": "known_answers", "kind": "hex",
                 "params": [k[1] for k in K.KNOWN], "caps": [k[2] for k in K.KNOWN],
                 "names": [k[0] for k in K.KNOWN]})
    for base in ("starfleet_dynamic", "starfleet_fixed"):
        comp, plain = bases[base]
        fams.append({"name": f"cut7_{base}", "kind": "cut", "base": base,
                     "params": list(range(0, len(comp), 7)) + [len(comp)],
                     "caps": len(plain)})
    for base in ("dynamic_4096", "fixed_4096", "stored_4096", "multiblock_12000"):
        comp, plain = bases[base]
        fams.append({"name": f"cut1_{base}", "kind": "cut", "base": base,
                     "params": list(range(0, min(len(comp), 400))) +
                     list(range(400, len(comp) + 1, 13)), "caps": len(plain)})
        nbits = len(comp) * 8
        flips = sorted(set(int(x) for x in rng.integers(0, nbits, 700)) |
                       set(range(0, min(nbits, 900))))
        fams.append({"name": f"flip_{base}", "kind": "flip", "base": base, "params": flips,
                     "caps": len(plain) + 64})
        fams.append({"name": f"cap_{base}", "kind": "cap", "base": base,
                     "params": sorted(set(list(range(0, 300)) +
                                          [int(x) for x in rng.integers(0, len(plain), 200)] +
                                          [len(plain) - 1, len(plain), len(plain) + 1, len(plain) + 100]))})
    comp, plain = bases["repetitive_70000"]
    fams.append({"name": "cap_repetitive_70000", "kind": "cap", "base": "repetitive_70000",
                 "params": sorted(set([int(x) for x in rng.integers(0, len(plain), 150)] +
                                      [0, 1, 257, 258, 259, len(plain) - 1, len(plain)]))})
    fams.append({"name": "flip_repetitive_70000", "kind": "flip", "base": "repetitive_70000",
                 "params": sorted(set(int(x) for x in rng.integers(0, len(comp) * 8, 400))),
                 "caps": len(plain)})
    crafted = crafted_cases(rng, 2400)
    fams.append({"name": "crafted_dynamic_headers", "kind": "hex",
                 "params": [c[0].hex() for c in crafted], "caps": [c[1] for c in crafted]})
    fuzz = []
    for _ in range(1500):
        b = bytearray(rng.integers(0, 256, int(rng.integers(1, 80)), dtype=np.uint8).tobytes())
        b[0] = (b[0] & 0xF8) | int(rng.choice([1, 2, 3, 4, 5, 0]))
        fuzz.append(bytes(b))
    fams.append({"name": "random_bytes", "kind": "hex", "params": [f.hex() for f in fuzz],
                 "caps": [int(rng.choice([0, 16, 300])) for _ in fuzz]})
    # the C2 unit itself (one 64 KiB level-6 text stream): truncations, bit flips, capacities
    comp, plain = bases["dynamic_65536"]
    fams.append({"name": "cut_dynamic_65536", "kind": "cut", "base": "dynamic_65536",
                 "params": list(range(0, 200)) + list(range(200, len(comp) + 1, 89)) + [len(comp) - 1, len(comp)],
                 "caps": len(plain)})
    nbits = len(comp) * 8
    fams.append({"name": "flip_dynamic_65536", "kind": "flip", "base": "dynamic_65536",
                 "params": sorted(set(int(x) for x in rng.integers(0, nbits, 500)) | set(range(0, 700))),
                 "caps": len(plain) + 64})
    fams.append({"name": "cap_dynamic_65536", "kind": "cap", "base": "dynamic_65536",
                 "params": sorted(set([int(x) for x in rng.integers(0, len(plain), 120)] +
                                      [0, 1, 2, 3, len(plain) - 1, len(plain), len(plain) + 1]))})
    return fams


def expand(fam, bases):
    """-> list of (src bytes, cap)"""
    kind = fam["kind"]
    if kind == "hex":
        return [(bytes.fromhex(h), c) for h, c in zip(fam["params"], fam["caps"])]
    comp, plain = bases[fam["base"]]
    if kind == "cut":
        return [(comp[:n], fam["caps"]) for n in fam["params"]]
    if kind == "flip":
        out = []
        for bit in fam["params"]:
            b = bytearray(comp)
            b[bit >> 3] ^= 1 << (bit & 7)
            out.append((bytes(b), fam["caps"]))
        return out
    if kind == "cap":
        return [(comp, c) for c in fam["params"]]
    raise ValueError(kind)


def run_classifier(exe, cases, timeout_s=5, jobs=8):
    """Run one classifier build over `cases`, split over `jobs` concurrent processes."""
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0",
               UBSAN_OPTIONS="halt_on_error=1")
    chunks = [cases[i::jobs] for i in range(jobs)]
    procs = []
    for chunk in chunks:
        with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
            f.write(struct.pack("<I", len(chunk)))
            for src, cap in chunk:
                f.write(struct.pack("<II", len(src), cap))
                f.write(src)
            path = f.name
        procs.append((path, subprocess.Popen(
            [os.path.join(ROOT, "oracle", "_ref", exe), path, str(timeout_s)],
            stdout=subprocess.PIPE, text=True, env=env)))
    res = [None] * len(cases)
    for j, (path, p) in enumerate(procs):
        out, _ = p.communicate()
        assert p.returncode == 0, (exe, p.returncode)
        os.unlink(path)
        lines = [l for l in out.split("\n") if l.strip()]
        assert len(lines) == len(chunks[j]), (exe, len(lines), len(chunks[j]))
        for k, line in enumerate(lines):
            idx, cls, st, h = line.split()
            res[j + k * jobs] = (cls, int(st), h)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", action="store_true", help="write tests/golden fixtures")
    args = ap.parse_args()
    bindings.build(with_reference=True)
    orc = bindings.load_oracle()
    bases = build_bases()
    fams = build_families(bases)
    totals = {"cases": 0, "defined": 0, "assert_only": 0, "undefined": 0}
    bad = 0
    for fam in fams:
        cases = expand(fam, bases)
        nd = run_classifier("classify_ndebug", cases)
        sa = run_classifier("classify_san", cases, timeout_s=60)
        asr = run_classifier("classify_assert", cases)
        o_status, o_written, o_ub, o_hash = [], [], [], []
        for src, cap in cases:
            st, dst, wr, ub = orc.decompress(src, cap)
            o_status.append(st); o_written.append(int(wr)); o_ub.append(int(ub))
            o_hash.append("%016x" % orc.fnv1a64(dst))
        cls = []
        for i, (src, cap) in enumerate(cases):
            if sa[i][0] == "R":
                c = "D" if asr[i][0] == "R" else "A"
                ok = (o_ub[i] == 0 and o_status[i] == sa[i][1] and o_hash[i] == sa[i][2]
                      and nd[i][0] == "R" and nd[i][1] == sa[i][1] and nd[i][2] == sa[i][2])
                if c == "D":
                    ok = ok and asr[i][1] == sa[i][1] and asr[i][2] == sa[i][2]
            else:
                c = "U"
                ok = o_ub[i] == 1
            if not ok:
                bad += 1
                if bad <= 20:
                    print(f"MISMATCH {fam['name']}[{i}] src={src[:40].hex()}.. cap={cap} san={sa[i]} "
                          f"ndebug={nd[i]} assert={asr[i]} oracle=({o_status[i]},{o_written[i]},"
                          f"ub={o_ub[i]},{o_hash[i]})")
            cls.append(c)
            totals["cases"] += 1
            totals[{"D": "defined", "A": "assert_only", "U": "undefined"}[c]] += 1
        fam["class"] = "".join(cls)
        fam["status"] = o_status
        fam["written"] = o_written
        fam["hash"] = o_hash
        # where defined, these ARE the reference's results (checked equal above)
        print(f"{fam['name']:32s} n={len(cases):5d}  D={cls.count('D')} A={cls.count('A')} U={cls.count('U')}")
    # the hand-written table must agree with what the binaries say
    ka = fams[0]
    for i, k in enumerate(K.KNOWN):
        name, hx, cap, st, prefix, c = k
        if ka["class"][i] != c or ka["status"][i] != st:
            bad += 1
            print(f"KNOWN-ANSWER TABLE MISMATCH {name}: table=({st},{c}) measured=({ka['status'][i]},{ka['class'][i]})")
        if prefix is not None:
            _, dst, wr, _ = orc.decompress(bytes.fromhex(hx), cap)
            if dst[:wr].hex() != prefix:
                bad += 1
                print(f"KNOWN-ANSWER PREFIX MISMATCH {name}: {dst[:wr].hex()} != {prefix}")
    print(totals, "mismatches:", bad)
    if bad:
        sys.exit(1)
    if args.write:
        os.makedirs(os.path.join(GOLDEN, "bases"), exist_ok=True)
        meta = {}
        for name, (comp, plain) in bases.items():
            with open(os.path.join(GOLDEN, "bases", name + ".deflate"), "wb") as f:
                f.write(comp)
            meta[name] = {"comp_len": len(comp), "plain_len": len(plain),
                          "plain_md5": hashlib.md5(plain).hexdigest(),
                          "plain_fnv1a64": "%016x" % orc.fnv1a64(plain)}
        with open(os.path.join(GOLDEN, "families.json"), "w") as f:
            json.dump({"generator": "oracle/pin_oracle.py", "zlib": zlib.ZLIB_RUNTIME_VERSION,
                       "fill": 0xA5, "totals": totals, "bases": meta, "families": fams}, f,
                      separators=(",", ":"))
        print("wrote", GOLDEN)


if __name__ == "__main__":
    main()
