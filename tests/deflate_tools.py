"""Test-side tooling: a bit-level DEFLATE stream writer, zlib wrappers and the
synthetic corpora named in BASELINE.json (C1..C5 units).  Harness only.

Bit-packing rule (RFC 1951 §3.1.1): header fields, extra bits and
stored bytes are written LSB-first; Huffman codes are written MSB-first; the
stream is zero-padded to a byte.
"""
from __future__ import annotations

import zlib

import numpy as np

CL_ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
            67, 83, 99, 115, 131, 163, 195, 227, 258]
LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3,
             4, 4, 4, 4, 5, 5, 5, 5, 0]
DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
             1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577]
DIST_EXTRA = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,
              9, 9, 10, 10, 11, 11, 12, 12, 13, 13]


def canonical_codes(lens):
    """RFC 1951 §3.2.2 code assignment. Returns {symbol: (code, bitsize)} for bitsize > 0.
    For over-subscribed sets the arithmetic simply continues (values may reach 2**bitsize)."""
    out = {}
    code = 0
    prev = 0
    for L in range(1, 16):
        syms = [s for s, l in enumerate(lens) if l == L]
        if not syms:
            continue
        code <<= (L - prev)
        prev = L
        for s in syms:
            out[s] = (code, L)
            code += 1
    return out


class BitWriter:
    def __init__(self):
        self.bits = []

    def field(self, value: int, n: int):
        """n-bit little-endian field (LSB first)."""
        for i in range(n):
            self.bits.append((value >> i) & 1)
        return self

    def code(self, value: int, n: int):
        """n-bit Huffman code (MSB first)."""
        for i in reversed(range(n)):
            self.bits.append((value >> i) & 1)
        return self

    def align(self):
        while len(self.bits) % 8:
            self.bits.append(0)
        return self

    def raw(self, data: bytes):
        assert len(self.bits) % 8 == 0
        for b in data:
            self.field(b, 8)
        return self

    def header(self, final: bool, btype: int):
        self.field(1 if final else 0, 1)
        self.field(btype, 2)
        return self

    def stored(self, final: bool, payload: bytes, nlen=None):
        self.header(final, 0).align()
        n = len(payload)
        self.field(n, 16).field((~n & 0xFFFF) if nlen is None else nlen, 16)
        return self.raw(payload)

    def tobytes(self) -> bytes:
        bits = list(self.bits)
        while len(bits) % 8:
            bits.append(0)
        out = bytearray()
        for i in range(0, len(bits), 8):
            v = 0
            for j in range(8):
                v |= bits[i + j] << j
            out.append(v)
        return bytes(out)

    def __len__(self):
        return len(self.bits)


FIXED_LIT_LENS = [8] * 144 + [9] * 112 + [7] * 24 + [8] * 8
FIXED_DIST_LENS = [5] * 32


class HuffBlock:
    """Writes symbols of one Huffman block given explicit code-length sets."""

    def __init__(self, w: BitWriter, lit_lens, dist_lens):
        self.w = w
        self.lit = canonical_codes(lit_lens)
        self.dist = canonical_codes(dist_lens)

    def lit_sym(self, s: int):
        c, n = self.lit[s]
        self.w.code(c, n)
        return self

    def literal(self, byte: int):
        return self.lit_sym(byte)

    def eob(self):
        return self.lit_sym(256)

    def length(self, length: int):
        idx = 28 if length == 258 else max(i for i in range(28) if LEN_BASE[i] <= length)
        self.lit_sym(257 + idx)
        self.w.field(length - LEN_BASE[idx], LEN_EXTRA[idx])
        return self

    def distance(self, dist: int):
        idx = max(i for i in range(30) if DIST_BASE[i] <= dist)
        c, n = self.dist[idx]
        self.w.code(c, n)
        self.w.field(dist - DIST_BASE[idx], DIST_EXTRA[idx])
        return self

    def match(self, length: int, dist: int):
        return self.length(length).distance(dist)


def fixed_block(w: BitWriter, final: bool) -> HuffBlock:
    w.header(final, 1)
    return HuffBlock(w, FIXED_LIT_LENS, FIXED_DIST_LENS)


def dynamic_header(w: BitWriter, final: bool, lit_lens, dist_lens, cl_lens=None,
                   cl_symbols=None, hclen=19):
    """Emit a dynamic block header.

    lit_lens / dist_lens : code length per symbol (len(lit_lens) in 257..288, len(dist_lens) in 1..32)
    cl_lens   : 19 code-length-code lengths (default: every symbol 5 bits -> code value == symbol)
    cl_symbols: explicit list of (cl_symbol, extra_value) pairs for the two runs
                [(lit run), (dist run)]; default: one plain symbol per length (no repeats).
    """
    w.header(final, 2)
    if cl_lens is None:
        cl_lens = [5] * 19
    w.field(len(lit_lens) - 257, 5).field(len(dist_lens) - 1, 5).field(hclen - 4, 4)
    for i in range(hclen):
        w.field(cl_lens[CL_ORDER[i]], 3)
    cl = canonical_codes(cl_lens)
    extra_bits = {16: 2, 17: 3, 18: 7}
    if cl_symbols is None:
        cl_symbols = [[(l, 0) for l in lit_lens], [(l, 0) for l in dist_lens]]
    for run in cl_symbols:
        for sym, extra in run:
            c, n = cl[sym]
            w.code(c, n)
            if sym in extra_bits:
                w.field(extra, extra_bits[sym])
    return HuffBlock(w, lit_lens, dist_lens)


# --------------------------------------------------------------------------- zlib side
def raw_deflate(data: bytes, level: int = 6, strategy: int = zlib.Z_DEFAULT_STRATEGY,
                mem_level: int = 8) -> bytes:
    """Same call shape as /root/reference/tools/deflate_compress.py:9-13 (raw DEFLATE)."""
    c = zlib.compressobj(level, zlib.DEFLATED, -15, mem_level, strategy)
    return c.compress(data) + c.flush()


def raw_deflate_multiblock(chunks, levels, strategies) -> bytes:
    """One stream whose blocks change type: Z_FULL_FLUSH between differently-compressed chunks
    is not expressible with one compressobj, so build it from stored/fixed/dynamic pieces by
    concatenating non-final compressed pieces ending on a byte boundary (Z_SYNC_FLUSH)."""
    out = bytearray()
    for i, (chunk, lvl, strat) in enumerate(zip(chunks, levels, strategies)):
        c = zlib.compressobj(lvl, zlib.DEFLATED, -15, 8, strat)
        last = i == len(chunks) - 1
        if last:
            out += c.compress(chunk) + c.flush()
        else:
            # sync flush terminates with an empty stored block (non-final) on a byte boundary
            out += c.compress(chunk) + c.flush(zlib.Z_SYNC_FLUSH)
    return bytes(out)


def first_block_type(stream: bytes) -> int:
    return (stream[0] >> 1) & 3


# --------------------------------------------------------------------------- synthetic corpora
_WORD_CHARS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)


def _vocab(rng: np.random.Generator, n_words: int):
    lens = np.clip(rng.poisson(4.5, n_words) + 1, 1, 14)
    probs = np.arange(1, 27, dtype=np.float64) ** -0.9
    probs /= probs.sum()
    words = []
    for L in lens:
        words.append(_WORD_CHARS[rng.choice(26, size=int(L), p=probs)].tobytes())
    return words


_VOCAB_CACHE = {}


def text_like(n_bytes: int, seed: int, vocab_size: int = 8192) -> bytes:
    """Seeded Zipf-distributed word soup with punctuation and light markup."""
    key = vocab_size
    if key not in _VOCAB_CACHE:
        vr = np.random.default_rng(12345)
        words = _vocab(vr, vocab_size)
        tags = [b"<p>", b"</p>", b"<a href=\"", b"\">", b"</a>", b"<div class=\"", b"</div>",
                b"<li>", b"</li>", b"<br/>"]
        punct = [b". ", b", ", b"; ", b"\n", b" - ", b"? ", b": "]
        table = words + tags + punct
        joined = np.frombuffer(b"".join(t if t in tags or t in punct else t + b" " for t in table),
                               dtype=np.uint8)
        lens = np.array([len(t) + (0 if (t in tags or t in punct) else 1) for t in table])
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
        ranks = np.arange(1, len(table) + 1, dtype=np.float64)
        # words Zipfian; markup/punctuation get fixed mid-range mass
        p = ranks ** -1.05
        p[len(words):] = p[40]
        p /= p.sum()
        _VOCAB_CACHE[key] = (joined, lens, offs, np.cumsum(p))
    joined, lens, offs, cdf = _VOCAB_CACHE[key]
    rng = np.random.default_rng(seed)
    n_tok = n_bytes // 4 + 64
    idx = np.searchsorted(cdf, rng.random(n_tok))
    idx = np.minimum(idx, len(lens) - 1)
    tl = lens[idx]
    ends = np.cumsum(tl)
    k = int(np.searchsorted(ends, n_bytes)) + 1
    idx, tl, ends = idx[:k], tl[:k], ends[:k]
    total = int(ends[-1])
    starts = ends - tl
    # gather: position j belongs to token t(j); source = offs[idx[t]] + (j - starts[t])
    tok_of = np.repeat(np.arange(k), tl)
    src = offs[idx][tok_of] + (np.arange(total) - starts[tok_of])
    out = joined[src][:n_bytes]
    if len(out) < n_bytes:  # extremely unlikely; pad with spaces
        out = np.concatenate([out, np.full(n_bytes - len(out), 32, np.uint8)])
    return out.tobytes()


def big_text(n_bytes: int, seed: int) -> bytes:
    """Text for one LARGE stream (BASELINE configs C1 / C5): distinct 4 MiB pieces of text_like up
    to 32 MiB, then repeated — a repeat lies far outside DEFLATE's 32 KiB window, so it compresses
    and decodes exactly like fresh text (and generating 1 GiB stays a matter of seconds)."""
    piece = 4 << 20
    unique = min(n_bytes, 32 << 20)
    out = b"".join(text_like(min(piece, unique - o), seed + o // piece) for o in range(0, unique, piece))
    return (out * ((n_bytes + unique - 1) // unique))[:n_bytes] if unique else b""


def repetitive(n_bytes: int, seed: int) -> bytes:
    """Long single-byte runs (distance-1, length-258 matches), short-period patterns and far
    repeats up to 32 KiB."""
    rng = np.random.default_rng(seed)
    out = bytearray()
    history = bytearray(rng.integers(0, 256, 512, dtype=np.uint8).tobytes())
    out += history
    while len(out) < n_bytes:
        kind = rng.random()
        if kind < 0.55:      # single byte run
            out += bytes([int(rng.integers(0, 256))]) * int(rng.integers(2000, 40000))
        elif kind < 0.8:     # short period pattern
            period = int(rng.integers(2, 24))
            pat = rng.integers(0, 256, period, dtype=np.uint8).tobytes()
            out += pat * (int(rng.integers(2000, 20000)) // period)
        else:                # far repeat from up to 32 KiB back
            back = int(rng.integers(1024, min(32768, len(out)) + 1)) if len(out) > 1024 else len(out)
            n = int(rng.integers(500, 8000))
            start = len(out) - back
            out += out[start:start + min(n, back)]
    return bytes(out[:n_bytes])


def incompressible(n_bytes: int, seed: int) -> bytes:
    return np.random.default_rng(seed).integers(0, 256, n_bytes, dtype=np.uint8).tobytes()


def make_stream(kind: str, n_bytes: int, seed: int):
    """-> (plain, compressed).  kind in dynamic|fixed|stored|repetitive|multiblock."""
    if kind == "dynamic":
        plain = text_like(n_bytes, seed)
        return plain, raw_deflate(plain, 6)
    if kind == "fixed":
        plain = text_like(n_bytes, seed)
        return plain, raw_deflate(plain, 6, zlib.Z_FIXED)
    if kind == "stored":
        plain = text_like(n_bytes, seed)
        return plain, raw_deflate(plain, 0)
    if kind == "repetitive":
        plain = repetitive(n_bytes, seed)
        return plain, raw_deflate(plain, 9)
    if kind == "multiblock":
        plain = text_like(n_bytes, seed)
        a, b = n_bytes // 3, 2 * n_bytes // 3
        comp = raw_deflate_multiblock([plain[:a], plain[a:b], plain[b:]], [6, 0, 6],
                                      [zlib.Z_FIXED, zlib.Z_DEFAULT_STRATEGY, zlib.Z_DEFAULT_STRATEGY])
        return plain, comp
    raise ValueError(kind)


# --------------------------------------------------------------------------- batches
class Batch:
    """Flat layout shared by the C-ABI, the oracle and the reference shim:
    stream i = src[src_off[i] : src_off[i]+src_len[i]] -> dst[dst_off[i] : dst_off[i]+dst_cap[i]]."""

    def __init__(self, streams, caps, dst_align: int = 1, src_align: int = 1):
        n = len(streams)
        self.n = n
        self.src_len = np.array([len(s) for s in streams], dtype=np.uint64)
        self.dst_cap = np.array(caps, dtype=np.uint64)

        def layout(sizes, align):
            off = np.zeros(n, dtype=np.uint64)
            at = 0
            for i, s in enumerate(sizes):
                at = (at + align - 1) // align * align
                off[i] = at
                at += int(s)
            return off, at

        self.src_off, src_total = layout(self.src_len, src_align)
        self.dst_off, dst_total = layout(self.dst_cap, dst_align)
        self.src = np.zeros(max(src_total, 1), dtype=np.uint8)
        for i, s in enumerate(streams):
            if len(s):
                o = int(self.src_off[i])
                self.src[o:o + len(s)] = np.frombuffer(s, dtype=np.uint8)
        self.dst_total = max(dst_total, 1)

    def new_dst(self, fill: int = 0xA5):
        return np.full(self.dst_total, fill, dtype=np.uint8)

    def dst_slice(self, dst, i):
        o = int(self.dst_off[i])
        return dst[o:o + int(self.dst_cap[i])]


def fnv1a64(data: bytes) -> int:
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


# --------------------------------------------------------------------------- crafted LZ77 shapes
def crafted_lz_streams():
    """Fixed-Huffman streams that walk the LZ77 resolve pass through its special cases at every
    phase of its 128-byte chunks: distance-1 runs, short periods, far non-overlapping copies,
    maximum-length matches back to back, matches that straddle chunk and word boundaries,
    matches whose source lies in the same chunk.  -> list of (name, compressed bytes, plain length)."""
    out = []
    rng = np.random.default_rng(99)
    for lead in (0, 1, 2, 3, 5, 31, 61, 126, 127, 128, 129, 200, 257):
        for dist, reps in ((1, 5), (2, 3), (3, 3), (4, 3), (5, 4), (6, 3), (7, 3), (9, 3), (13, 2), (24, 2), (127, 2),
                           (128, 3), (129, 3), (200, 3), (258, 2), (1000, 2)):
            w = BitWriter()
            blk = fixed_block(w, True)
            n = max(lead, dist)
            for b in rng.integers(32, 127, n):
                blk.literal(int(b))
            total = n
            for r in range(reps):
                ln = 258 if r % 2 == 0 else int(rng.integers(3, 258))
                blk.match(ln, dist)
                total += ln
                blk.literal(int(rng.integers(32, 127)))
                total += 1
            blk.eob()
            out.append((f"lead{lead}_d{dist}", w.tobytes(), total))
    # dense short matches: every 3-5 bytes a match, near and far sources mixed
    w = BitWriter()
    blk = fixed_block(w, True)
    for b in rng.integers(32, 127, 300):
        blk.literal(int(b))
    total = 300
    for i in range(400):
        d = int(rng.choice([1, 2, 3, 4, 5, 8, 16, 33, 64, 100, 127, 128, 129, 250, 299]))
        ln = int(rng.choice([3, 3, 4, 5, 6, 9, 17, 40]))
        blk.match(ln, min(d, total))
        total += ln
        if i % 3 == 0:
            blk.literal(int(rng.integers(32, 127)))
            total += 1
    blk.eob()
    out.append(("dense_short", w.tobytes(), total))
    # back-to-back 3-byte matches (two heads in one 4-byte word) at every word phase, sources near
    # the very start of the stream (the word below the first byte must not be read)
    for lead in (1, 2, 3, 4, 5, 6, 7):
        w = BitWriter()
        blk = fixed_block(w, True)
        for b in rng.integers(32, 127, lead):
            blk.literal(int(b))
        total = lead
        for i in range(120):
            d = min(total, int(rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 30, 131])))
            if i % 5 == 4:
                d = total  # the source is the first bytes of the stream
            blk.match(3 if i % 3 else 4, d)
            total += 3 if i % 3 else 4
        blk.eob()
        out.append((f"threes_lead{lead}", w.tobytes(), total))
    # RUNS: back-to-back maximum-length matches with one distance (what a compressor makes of a long
    # repeat), a shorter last one, then another run with another distance right behind — the resolve
    # pass merges such runs and fills them 16 bytes per lane; every kind of period, every phase
    for lead in (0, 1, 7, 15, 16, 17, 100, 1000, 1023, 1024, 1025):
        for dist, nmax, last in ((1, 3, 100), (1, 40, 0), (2, 5, 3), (3, 4, 257), (4, 2, 50), (5, 3, 0), (8, 33, 258),
                                 (15, 2, 9), (16, 6, 77), (17, 3, 30), (31, 2, 0), (258, 3, 11), (300, 4, 100),
                                 (511, 2, 3), (512, 5, 200), (513, 35, 17), (4000, 3, 0),
                                 (6, 3, 5), (7, 4, 0), (9, 3, 40), (11, 2, 3), (13, 3, 99), (23, 2, 7), (100, 3, 0)):
            w = BitWriter()
            blk = fixed_block(w, True)
            n = max(lead, dist)
            for b in rng.integers(32, 127, n):
                blk.literal(int(b))
            total = n
            for d2, k, l2 in ((dist, nmax, last), (max(1, dist // 2), 2, 5)):
                for r in range(k):
                    blk.match(258, d2)
                    total += 258
                if l2 >= 3:
                    blk.match(l2, d2)
                    total += l2
            blk.literal(65)
            total += 1
            blk.eob()
            out.append((f"run_lead{lead}_d{dist}x{nmax}", w.tobytes(), total))
    return out
