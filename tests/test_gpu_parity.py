"""GPU parity: the CUDA path, called through the C ABI, against the golden fixtures (the
reference's own results), the oracle on seeded batches, and size-independent properties at
BASELINE.json's full sizes.  Integer/byte work: the bar is bit-exact (bytes, status, written)."""
import hashlib

import zlib

import os

import numpy as np
import pytest
import torch

from tests import deflate_tools as T
from tests import gpu_util
from tests import known_answers as K

pytestmark = pytest.mark.gpu


def _check_family(ctx, oracle, golden, name, dst_align=1, src_align=1):
    cases = golden.cases(name)
    b = T.Batch([c[1] for c in cases], [c[2] for c in cases], dst_align=dst_align,
                src_align=src_align)
    st, wr, dst = gpu_util.run_device(ctx, b)
    bad = []
    for k, (i, src, cap) in enumerate(cases):
        want_st, want_wr, want_hash, cls = golden.expected(name, i)
        got_hash = "%016x" % oracle.fnv1a64(b.dst_slice(dst, k).tobytes())
        if (int(st[k]), int(wr[k]), got_hash) != (want_st, want_wr, want_hash):
            bad.append((i, cls, int(st[k]), int(wr[k]), want_st, want_wr))
    assert not bad, (name, len(bad), bad[:5])
    # nothing outside the per-stream regions may change (gaps keep the fill value)
    mask = np.ones(b.dst_total, dtype=bool)
    for k in range(b.n):
        o = int(b.dst_off[k])
        mask[o:o + int(b.dst_cap[k])] = False
    assert (dst[mask] == 0xA5).all()


def test_golden_families_bit_exact(ctx, oracle, golden):
    """All 25k committed cases: status, written and every dst byte (incl. untouched tail)."""
    for name in golden.families:
        _check_family(ctx, oracle, golden, name)


@pytest.mark.parametrize("dst_align,src_align", [(8, 4), (256, 16), (3, 5)])
def test_golden_alignment_variants(ctx, oracle, golden, dst_align, src_align):
    for name in ("known_answers", "cut1_multiblock_12000", "cap_dynamic_4096",
                 "cap_stored_4096", "crafted_dynamic_headers", "flip_repetitive_70000"):
        _check_family(ctx, oracle, golden, name, dst_align, src_align)


@pytest.mark.parametrize("name,hx,cap,status,prefix,cls", K.KNOWN, ids=[k[0] for k in K.KNOWN])
def test_known_answers_single_stream_host_api(ctx, name, hx, cap, status, prefix, cls):
    """sfb200_decompress: the reference entry point's shape, host buffers."""
    st, dst, written = ctx.decompress(bytes.fromhex(hx), cap, fill=0xA5)
    assert st == status
    if prefix is not None:
        assert dst[:written].hex() == prefix
    assert dst[written:] == bytes([0xA5]) * (cap - written)


def test_starfleet_roundtrip(ctx, golden):
    # /root/reference/src/test/decompress_test.cpp:136-174
    for base in ("starfleet_fixed", "starfleet_dynamic"):
        st, dst, written = ctx.decompress(golden.bases[base], K.STARFLEET_LEN)
        assert (st, written) == (0, K.STARFLEET_LEN)
        assert hashlib.md5(dst).hexdigest() == K.STARFLEET_MD5


def test_mixed_batch_vs_oracle(ctx, oracle):
    """Seeded ragged batch (all block types, truncations, short dst) — every byte vs the oracle."""
    rng = np.random.default_rng(77)
    streams, caps = [], []
    kinds = ["dynamic", "fixed", "stored", "multiblock", "repetitive"]
    for i in range(1500):
        size = int(rng.choice([1, 17, 300, 4096, 20000, 65536, 150000]))
        plain, comp = T.make_stream(kinds[i % 5], size, 5000 + i)
        if i % 11 == 0:
            comp = comp[: int(rng.integers(0, len(comp) + 1))]
        cap = len(plain)
        if i % 13 == 0:
            cap = max(0, cap - int(rng.integers(1, 300)))
        if i % 17 == 0:
            cap += int(rng.integers(1, 50))
        streams.append(comp)
        caps.append(cap)
    streams.append(b"")  # empty source
    caps.append(0)
    b = T.Batch(streams, caps)
    st, wr, dst = gpu_util.run_device(ctx, b)
    dst_o = b.new_dst()
    ost, owr, _ = oracle.decompress_batch(b.src, b.src_off, b.src_len, dst_o, b.dst_off, b.dst_cap, 8)
    assert (st == ost).all(), np.nonzero(st != ost)[0][:10]
    assert (wr == owr).all()
    assert (dst == dst_o).all()
    assert set(np.unique(ost)) >= {0, 4, 6}  # the batch really contains failures too


def test_host_batch_api_matches_device_api(ctx, oracle):
    streams, caps = [], []
    for i in range(200):
        plain, comp = T.make_stream(["dynamic", "fixed", "stored"][i % 3], 100 + 37 * i, i)
        streams.append(comp)
        caps.append(len(plain) - (i % 5 == 0) * 3 + (i % 7 == 0) * 9)
    b = T.Batch(streams, caps, dst_align=1)
    dst = b.new_dst()
    st, wr = ctx.decompress_batch_host(b.src, b.src_off, b.src_len, dst, b.dst_off, b.dst_cap)
    dst_o = b.new_dst()
    ost, owr, _ = oracle.decompress_batch(b.src, b.src_off, b.src_len, dst_o, b.dst_off, b.dst_cap)
    assert (st == ost).all() and (wr == owr).all() and (dst == dst_o).all()


def test_host_batch_api_pipelined_sub_batches(ctx, oracle, monkeypatch):
    """The host entry point cut into many sub-batches (H2D / kernels / D2H overlapped): ragged
    regions, gaps, failures and short capacities must come back exactly as from one batch, and
    every byte the decoder did not produce must keep the caller's value."""
    monkeypatch.setenv("SFB200_HOST_STAGING_KB", "48")
    rng = np.random.default_rng(3)
    streams, caps = [], []
    for i in range(300):
        kind = ["dynamic", "fixed", "stored", "multiblock", "repetitive"][i % 5]
        plain, comp = T.make_stream(kind, int(rng.integers(1, 30000)), 9000 + i)
        if i % 9 == 0:
            comp = comp[: int(rng.integers(0, len(comp)))]
        streams.append(comp)
        caps.append(max(0, len(plain) + int(rng.integers(-40, 40)) if i % 4 == 0 else len(plain)))
    for dst_align in (1, 64):
        b = T.Batch(streams, caps, dst_align=dst_align)
        dst = b.new_dst()
        st, wr = ctx.decompress_batch_host(b.src, b.src_off, b.src_len, dst, b.dst_off, b.dst_cap)
        dst_o = b.new_dst()
        ost, owr, _ = oracle.decompress_batch(b.src, b.src_off, b.src_len, dst_o, b.dst_off, b.dst_cap)
        assert (st == ost).all() and (wr == owr).all() and (dst == dst_o).all()


def test_host_batch_larger_than_the_staging_budget(oracle, monkeypatch):
    """SURVEY.md §8 f3 (bounded device memory): a batch several times the configured staging budget
    decodes through the ring of three sub-batch slots, bit-exact, failures and short capacities
    included, and the context's staging never grows beyond the budget."""
    import starflate_b200 as S
    monkeypatch.setenv("SFB200_HOST_STAGING_MB", "6")
    c = S.Context(0)
    try:
        rng = np.random.default_rng(5)
        streams, caps = [], []
        for i in range(800):
            kind = ["dynamic", "fixed", "stored", "multiblock", "repetitive"][i % 5]
            plain, comp = T.make_stream(kind, int(rng.integers(20000, 70000)), 12000 + i)
            if i % 13 == 0:
                comp = comp[: int(rng.integers(0, len(comp)))]
            streams.append(comp)
            caps.append(len(plain) - (5 if i % 11 == 0 else 0))
        b = T.Batch(streams, caps, dst_align=1)
        assert b.dst_total > 5 * (6 << 20)
        dst = b.new_dst()
        st, wr = c.decompress_batch_host(b.src, b.src_off, b.src_len, dst, b.dst_off, b.dst_cap)
        dst_o = b.new_dst()
        ost, owr, _ = oracle.decompress_batch(b.src, b.src_off, b.src_len, dst_o, b.dst_off, b.dst_cap)
        assert (st == ost).all() and (wr == owr).all() and (dst == dst_o).all()
        assert c.staging_bytes() <= (6 << 20) + 3 * 1024
    finally:
        c.close()


def test_one_batch_across_several_contexts(oracle):
    """sfb200_decompress_batch_host_multi: one batch cut into contiguous shards, one per context
    (normally one per GPU; two contexts on the devices that are there), each on its own host
    thread — results as from one call, and the caller's current device is left alone."""
    import torch

    import starflate_b200 as S
    ndev = torch.cuda.device_count()
    ctxs = [S.Context(k % ndev) for k in range(max(2, min(ndev, 4)))]
    try:
        rng = np.random.default_rng(6)
        streams, caps = [], []
        for i in range(400):
            kind = ["dynamic", "fixed", "stored", "multiblock", "repetitive"][i % 5]
            plain, comp = T.make_stream(kind, int(rng.integers(1, 40000)), 15000 + i)
            streams.append(comp if i % 10 else comp[: len(comp) // 3])
            caps.append(len(plain) - (3 if i % 6 == 0 else 0))
        b = T.Batch(streams, caps, dst_align=1)
        before = torch.cuda.current_device()
        dst = b.new_dst()
        st, wr = S.Context.decompress_batch_host_multi(ctxs, b.src, b.src_off, b.src_len, dst, b.dst_off, b.dst_cap)
        assert torch.cuda.current_device() == before
        dst_o = b.new_dst()
        ost, owr, _ = oracle.decompress_batch(b.src, b.src_off, b.src_len, dst_o, b.dst_off, b.dst_cap)
        assert (st == ost).all() and (wr == owr).all() and (dst == dst_o).all()
    finally:
        for c in ctxs:
            c.close()


def test_small_table_geometry_with_hand_over(oracle, golden, monkeypatch):
    """SFB200_GEOMETRY=small: pass 1 tries 6/5-bit-root tables (16 warps per SM) and hands the
    streams whose first block does not fit them (real HTML) to the large geometry."""
    import starflate_b200 as S
    monkeypatch.setenv("SFB200_GEOMETRY", "small")
    c = S.Context(0)
    try:
        for name in ("known_answers", "cut7_starfleet_dynamic", "cut1_multiblock_12000", "cap_dynamic_4096",
                     "crafted_dynamic_headers", "flip_repetitive_70000", "cap_stored_4096"):
            _check_family(c, oracle, golden, name)
    finally:
        c.close()


@pytest.mark.parametrize("env", [{"SFB200_NO_WIDE": "1"}, {"SFB200_NO_WIDE": "1", "SFB200_NO_STORED": "1", "SFB200_NO_PAIR": "1"},
                                 {"SFB200_NO_WIDE": "1", "SFB200_QUEUE": "1"}],
                         ids=["large-geometry", "large-geometry-plain", "queue-mode"])
def test_golden_families_on_the_large_geometry(oracle, golden, monkeypatch, env):
    """Small batches take the wide table geometry (9/6-bit roots, one CTA per SM) and stored streams
    the copy kernel: here every golden family goes through the 8/6-bit geometry the full-size
    batches use, once with and once without the copy kernel and the literal pairing, and once with
    pass 2 running beside pass 1 (the experimental hand-over queue: 16 KiB segments)."""
    import starflate_b200 as S
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    c = S.Context(0)
    try:
        for name in golden.families:
            _check_family(c, oracle, golden, name)
    finally:
        c.close()


def test_crafted_lz_shapes_and_unaligned_dst_base(ctx, oracle):
    """Runs, short periods, far copies and dense short matches at every chunk phase
    (tests.deflate_tools.crafted_lz_streams), packed back to back, through a dst pointer that is
    NOT 128-byte aligned (the C ABI re-bases it), incl. many 1-3 byte streams sharing bitmap words."""
    cases = T.crafted_lz_streams()
    streams = [c[1] for c in cases] + [bytes.fromhex("730400")] * 40 + [bytes.fromhex("731c0500")] * 3
    caps = [c[2] for c in cases] + [1] * 20 + [3] * 20 + [259, 258, 300]
    for dst_align in (1, 7):
        b = T.Batch(streams, caps, dst_align=dst_align)
        dst_o = b.new_dst()
        ost, owr, _ = oracle.decompress_batch(b.src, b.src_off, b.src_len, dst_o, b.dst_off, b.dst_cap)
        assert (ost[: len(cases)] == 0).all()
        dev = torch.device("cuda", ctx.device)
        for shift in (0, 13, 77):
            big = torch.full((b.dst_total + 256,), 0xA5, dtype=torch.uint8, device=dev)
            dst = big[shift: shift + b.dst_total]           # data_ptr() is base + shift
            t = lambda a: torch.from_numpy(a.view(np.int64)).to(dev)
            status = torch.full((b.n,), 0xEE, dtype=torch.uint8, device=dev)
            written = torch.full((b.n,), -1, dtype=torch.int64, device=dev)
            ctx.decompress_batch_device(torch.from_numpy(b.src).to(dev), t(b.src_off), t(b.src_len), dst,
                                        t(b.dst_off), t(b.dst_cap), status, written)
            torch.cuda.synchronize(dev)
            assert (status.cpu().numpy() == ost).all()
            assert (written.cpu().numpy().view(np.uint64) == owr).all()
            got = big.cpu().numpy()
            assert (got[shift: shift + b.dst_total] == dst_o).all()
            assert (got[:shift] == 0xA5).all() and (got[shift + b.dst_total:] == 0xA5).all()


@pytest.mark.parametrize("mode", ["1", "0"])
def test_single_stream_mode_and_batch_mode_agree_with_golden(oracle, golden, monkeypatch, mode):
    """SFB200_STREAM_MODE=1 forces the speculative one-warp-per-stream pass 1 (huff_stream.cuh) on
    batches of any size, =0 forces the lane-per-stream pass 1 even on a handful of streams: both
    must reproduce the reference's answers."""
    import starflate_b200 as S
    monkeypatch.setenv("SFB200_STREAM_MODE", mode)
    c = S.Context(0)
    try:
        for name in ("known_answers", "cut7_starfleet_dynamic", "cut1_multiblock_12000", "cap_dynamic_4096",
                     "crafted_dynamic_headers", "flip_repetitive_70000", "cap_stored_4096", "cut1_fixed_4096"):
            if name in golden.families:
                _check_family(c, oracle, golden, name, dst_align=1 if mode == "1" else 8)
    finally:
        c.close()


def test_large_single_streams(ctx, oracle):
    """C1-like shapes: a few multi-megabyte multi-block streams (single-stream mode picks them up),
    every byte against the oracle."""
    streams, caps, plains = [], [], []
    for kind, size, seed in (("dynamic", 3 << 20, 1), ("multiblock", 1 << 20, 2), ("repetitive", 4 << 20, 3),
                             ("dynamic", 1 << 20, 4)):
        plain, comp = T.make_stream(kind, size, 8800 + seed)
        streams.append(comp)
        caps.append(len(plain))
        plains.append(plain)
    streams.append(streams[0][: len(streams[0]) // 2])   # truncated in the middle of a block
    caps.append(caps[0])
    streams.append(streams[3])                            # destination too small
    caps.append(caps[3] - 12345)
    b = T.Batch(streams, caps, dst_align=1)
    st, wr, dst = gpu_util.run_device(ctx, b)
    dst_o = b.new_dst()
    ost, owr, _ = oracle.decompress_batch(b.src, b.src_off, b.src_len, dst_o, b.dst_off, b.dst_cap, 8)
    assert (st == ost).all() and (wr == owr).all()
    assert (dst == dst_o).all()
    assert list(st[:4]) == [0, 0, 0, 0] and st[4] != 0 and st[5] == 4


def _one_stream_per_call(ctx, src_bytes, cap, phase=0, src_phase=0, fill=0xA5):
    """n == 1: the single-stream route (block finder, a warp per block, pointer-jumping pass 2).
    dst sits `phase` bytes into a padded device buffer; -> (status, written, dst bytes, padding intact)"""
    import torch
    dev = torch.device("cuda", ctx.device)
    pad = 256
    s = torch.frombuffer(bytearray(b"\xEE" * src_phase + src_bytes + b"\xEE" * 8), dtype=torch.uint8).to(dev)
    big = torch.full((pad + phase + cap + pad,), fill, dtype=torch.uint8, device=dev)
    one = lambda v: torch.tensor([v], dtype=torch.int64, device=dev)
    status = torch.full((1,), 0xEE, dtype=torch.uint8, device=dev)
    written = torch.full((1,), -1, dtype=torch.int64, device=dev)
    ctx.decompress_batch_device(s, one(src_phase), one(len(src_bytes)), big[pad + phase:], one(0), one(cap), status,
                                written)
    torch.cuda.synchronize(dev)
    got = big.cpu().numpy()
    intact = bool((got[:pad + phase] == fill).all() and (got[pad + phase + cap:] == fill).all())
    return int(status.item()), int(written.item()), got[pad + phase: pad + phase + cap].tobytes(), intact


def test_single_stream_route_golden_one_call_per_stream(oracle, golden, monkeypatch):
    """Every 9th..61st case of the golden families, one call each (so each goes through the block
    finder, the counting / chain / writing jobs and the pointer-jumping pass 2; SFB200_STREAM_MODE=1
    because on its own the library takes that route only from 32 KiB of dst up), at varying dst and
    src phases."""
    import starflate_b200 as S
    monkeypatch.setenv("SFB200_STREAM_MODE", "1")
    ctx = S.Context(0)
    k = 0
    for name, stride in (("known_answers", 1), ("crafted_dynamic_headers", 1), ("cut7_starfleet_dynamic", 17),
                         ("cut7_starfleet_fixed", 61), ("cut1_multiblock_12000", 61), ("cap_dynamic_4096", 31),
                         ("cap_stored_4096", 61), ("flip_dynamic_4096", 31), ("flip_repetitive_70000", 61),
                         ("cap_repetitive_70000", 997)):
        if name not in golden.families:
            continue
        for i, src, cap in golden.cases(name, stride):
            want_st, want_wr, want_hash, cls = golden.expected(name, i)
            st, wr, dst, intact = _one_stream_per_call(ctx, src, cap, phase=(7 * k) % 131, src_phase=k % 5)
            k += 1
            assert (st, wr) == (want_st, want_wr), (name, i, cls, st, wr)
            assert "%016x" % oracle.fnv1a64(dst) == want_hash, (name, i, cls)
            assert intact, (name, i)
    assert k > 600
    ctx.close()


@pytest.mark.parametrize("blocks,chain", [("1", "parallel"), ("1", "sequential"), ("0", "parallel")])
def test_single_stream_route_large(oracle, monkeypatch, blocks, chain):
    """C1-shaped inputs (about a megabyte and more, many dynamic blocks; also stored and fixed blocks in
    between, a truncated input, a destination that is too small), blocks side by side
    (SFB200_BLOCKS=1, the default) and front to back: all bytes against the oracle."""
    import starflate_b200 as S
    monkeypatch.setenv("SFB200_BLOCKS", blocks)
    monkeypatch.setenv("SFB200_CHAIN", chain)  # the chain of blocks by one block of threads | by one thread
    c = S.Context(0)
    try:
        inputs = []
        for kind, size, seed in (("dynamic", 3 << 20, 1), ("multiblock", 1 << 20, 2), ("repetitive", 4 << 20, 3)):
            plain, comp = T.make_stream(kind, size, 8800 + seed)
            inputs.append((comp, len(plain)))
        big_plain = b"".join(T.text_like(1 << 20, 50 + j) for j in range(6))
        big = T.raw_deflate(big_plain, 6)
        inputs += [(big, len(big_plain)), (big[: len(big) // 2], len(big_plain)), (big, len(big_plain) - 54321),
                   (big, 1000)]
        mixed = T.raw_deflate_multiblock([T.text_like(300000, 1), T.incompressible(70000, 2), T.text_like(200000, 3),
                                          T.text_like(150000, 4)], [6, 0, 6, 9],
                                         [zlib.Z_DEFAULT_STRATEGY, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED,
                                          zlib.Z_DEFAULT_STRATEGY])
        inputs.append((mixed, 720000))
        for k, (src, cap) in enumerate(inputs):
            ost, odst, owr, _ = oracle.decompress(src, cap)
            st, wr, dst, intact = _one_stream_per_call(c, src, cap, phase=(29 * k) % 128, src_phase=k % 4)
            assert (st, wr) == (ost, owr), (k, st, wr, ost, owr)
            assert dst == odst and intact, k
    finally:
        c.close()


def test_one_bad_stream_does_not_affect_neighbours(ctx, oracle):
    plain, comp = T.make_stream("dynamic", 30000, 1)
    streams = [comp, comp[:100], comp, b"\x07", comp]
    b = T.Batch(streams, [len(plain)] * 5)
    st, wr, dst = gpu_util.run_device(ctx, b)
    assert list(st) == [oracle.decompress(s, len(plain))[0] for s in streams]
    assert st[0] == 0 and st[1] != 0 and st[3] == 2
    for k in (0, 2, 4):
        assert b.dst_slice(dst, k).tobytes() == plain


def _replicated_batch(unique, copies, dev):
    """`copies` x the unique streams: repeated src offsets, distinct dst regions (full-size runs
    without generating hundreds of thousands of unique inputs in the test)."""
    plains = [u[0] for u in unique]
    b = T.Batch([u[1] for u in unique], [len(p) for p in plains], dst_align=1)
    n_u = len(unique)
    sel = np.tile(np.arange(n_u), copies)
    src_off = b.src_off[sel]
    src_len = b.src_len[sel]
    cap = b.dst_cap[sel]
    dst_off = np.concatenate([[0], np.cumsum(cap)[:-1]]).astype(np.uint64)
    total = int(cap.sum())
    t = lambda a: torch.from_numpy(a.view(np.int64)).to(dev)
    return (torch.from_numpy(b.src).to(dev), t(src_off), t(src_len), t(dst_off), t(cap), total, sel,
            [gpu_util.host_checksum(p) for p in plains])


@pytest.mark.parametrize("shape", ["C2", "C3", "C4"])
def test_full_size_checksum_of_checksums(ctx, shape):
    """BASELINE.json configs at full size; verified through the device-side position-weighted
    checksum of every output region against the checksum of the known plaintext."""
    dev = torch.device("cuda", ctx.device)
    if shape == "C2":    # 65,536 x 64 KiB dynamic level 6
        unique = [T.make_stream("dynamic", 65536, 100 + i) for i in range(256)]
        copies = 256
    elif shape == "C3":  # 1,048,576 x 4 KiB mixed stored/fixed/dynamic
        unique = [T.make_stream(["stored", "fixed", "dynamic"][i % 3], 4096, 300 + i)
                  for i in range(1536)]
        copies = 1048576 // 1536 + 1
    else:                # 16,384 x 1 MiB level 9 repetitive
        unique = [T.make_stream("repetitive", 1 << 20, 700 + i) for i in range(32)]
        copies = 512
    src, src_off, src_len, dst_off, cap, total, sel, sums = _replicated_batch(unique, copies, dev)
    n = src_off.numel()
    dst = torch.zeros(total + 64, dtype=torch.uint8, device=dev)
    status = torch.full((n,), 0xEE, dtype=torch.uint8, device=dev)
    written = torch.zeros(n, dtype=torch.int64, device=dev)
    ctx.decompress_batch_device(src, src_off, src_len, dst, dst_off, cap, status, written)
    out = torch.zeros(n, dtype=torch.int64, device=dev)
    ctx.checksum_batch_device(dst, dst_off, written, out)
    torch.cuda.synchronize(dev)
    assert int(status.max()) == 0
    assert torch.equal(written, cap)
    want = np.array(sums, dtype=np.uint64)[sel]
    got = out.cpu().numpy().view(np.uint64)
    assert (got == want).all()
    # idempotence: a second run over the same dst gives the same bytes
    ctx.decompress_batch_device(src, src_off, src_len, dst, dst_off, cap, status, written)
    ctx.checksum_batch_device(dst, dst_off, written, out)
    torch.cuda.synchronize(dev)
    assert (out.cpu().numpy().view(np.uint64) == want).all()
    if shape == "C3":
        # the wave form (pass 2 of wave k beside pass 1 of wave k+1 on two internal streams; an
        # experiment since the end of round 2, SFB200_OVERLAP=1) gives the same bytes
        import starflate_b200 as S
        os.environ["SFB200_OVERLAP"] = "1"
        try:
            c2 = S.Context(ctx.device)
        finally:
            del os.environ["SFB200_OVERLAP"]
        try:
            dst.fill_(0x5A)
            os.environ["SFB200_OVERLAP"] = "1"   # (read per call)
            c2.decompress_batch_device(src, src_off, src_len, dst, dst_off, cap, status, written)
            c2.checksum_batch_device(dst, dst_off, written, out)
            torch.cuda.synchronize(dev)
            assert int(status.max()) == 0 and (out.cpu().numpy().view(np.uint64) == want).all()
        finally:
            os.environ.pop("SFB200_OVERLAP", None)
            c2.close()


@pytest.mark.parametrize("shape,size", [("C1", 1 << 20), ("C5", 1 << 30)])
def test_full_size_single_stream_round_trip(ctx, shape, size):
    """BASELINE configs C1 (one ~1 MiB text stream) and C5 (one 1 GiB text stream, ~12 000 dynamic
    blocks) at full size through the device API: Success, written == size, and every byte equal
    to the text that was compressed (compress -> decompress round trip; zlib made the stream, so
    no oracle run is needed at this size), plus the position-weighted checksum of the output
    against the same checksum of the input computed by the same kernel."""
    plain = T.big_text(size, 777)
    comp = T.raw_deflate(plain, 6)
    dev = torch.device("cuda", ctx.device)
    src = torch.frombuffer(bytearray(comp), dtype=torch.uint8).to(dev)
    want = torch.frombuffer(bytearray(plain), dtype=torch.uint8).to(dev)
    del plain
    dst = torch.full((size + 100,), 0xA5, dtype=torch.uint8, device=dev)
    one = lambda v: torch.tensor([v], dtype=torch.int64, device=dev)
    status = torch.full((1,), 0xEE, dtype=torch.uint8, device=dev)
    written = torch.full((1,), -1, dtype=torch.int64, device=dev)
    ctx.decompress_batch_device(src, one(0), one(len(comp)), dst, one(0), one(size + 100), status, written)
    torch.cuda.synchronize(dev)
    assert (int(status.item()), int(written.item())) == (0, size)
    assert torch.equal(dst[:size], want)
    assert bool((dst[size:] == 0xA5).all())  # capacity beyond the decoded size is left alone
    sums = torch.zeros(2, dtype=torch.int64, device=dev)
    ctx.checksum_batch_device(dst, one(0), written, sums[:1])
    ctx.checksum_batch_device(want, one(0), written, sums[1:])
    torch.cuda.synchronize(dev)
    assert int(sums[0].item()) == int(sums[1].item())


@pytest.mark.parametrize("stripe_mb", ["0", "8"])
def test_single_stream_route_across_pass2_stripes(monkeypatch, stripe_mb):
    """Pass 2 of the single-stream route, in one piece (default) and in 8 MiB stripes
    (SFB200_JUMP_STRIPE_MB): matches that straddle a stripe boundary (every boundary lies inside a
    length-258 match here), chains as long as the output (a run of one byte), plain text."""
    import starflate_b200 as S
    monkeypatch.setenv("SFB200_JUMP_STRIPE_MB", stripe_mb)
    ctx = S.Context(0)
    size = 20 << 20
    inputs = [b"abcdefg" * (size // 7), b"\x00" * size, T.big_text(24 << 20, 31),
              (b"xy" * 70000 + T.text_like(100000, 5)) * (size // 240000)]
    for k, plain in enumerate(inputs):
        comp = T.raw_deflate(plain, 9 if k != 2 else 6)
        st, wr, dst, intact = _one_stream_per_call(ctx, comp, len(plain) + 7, phase=(53 * k) % 128, src_phase=k)
        assert (st, wr) == (0, len(plain)) and intact, (k, st, wr)
        assert dst[: len(plain)] == plain, k
        assert dst[len(plain):] == b"\xa5" * 7
    ctx.close()


def test_size_discovery_matches_decode(ctx, oracle, golden):
    """sfb200_decompressed_size_batch_device (extension): status and size of a decode
    into an unlimited dst, without a dst — against the oracle with a large dst on golden families
    (truncated, corrupt and crafted inputs included), and against the full decode on a C2-shaped
    batch; sizing dst from it then lets every stream succeed."""
    dev = torch.device("cuda", ctx.device)
    as_i64 = lambda a: torch.from_numpy(a.view(np.int64)).to(dev)
    for name, stride in (("known_answers", 1), ("crafted_dynamic_headers", 1), ("cut7_starfleet_dynamic", 7),
                         ("cut1_multiblock_12000", 11), ("cap_stored_4096", 37), ("flip_dynamic_4096", 5),
                         ("flip_repetitive_70000", 9), ("cut1_fixed_4096", 7)):
        if name not in golden.families:
            continue
        cases = golden.cases(name, stride)
        big = 1 << 18
        b = T.Batch([c[1] for c in cases], [big] * len(cases))
        status = torch.full((b.n,), 0xEE, dtype=torch.uint8, device=dev)
        size = torch.full((b.n,), -1, dtype=torch.int64, device=dev)
        ctx.decompressed_size_batch_device(torch.from_numpy(b.src).to(dev), as_i64(b.src_off), as_i64(b.src_len),
                                           status, size)
        torch.cuda.synchronize(dev)
        dst_o = b.new_dst()
        ost, owr, _ = oracle.decompress_batch(b.src, b.src_off, b.src_len, dst_o, b.dst_off, b.dst_cap, 8)
        assert (status.cpu().numpy() == ost).all(), name
        assert (size.cpu().numpy().view(np.uint64) == owr).all(), name
    # sizes first, then a decode into exactly that much room
    streams = [T.make_stream(["dynamic", "fixed", "stored", "multiblock", "repetitive"][i % 5], 100 + 977 * i, 40 + i)[1]
               for i in range(300)]
    b0 = T.Batch(streams, [0] * len(streams))
    status = torch.zeros(b0.n, dtype=torch.uint8, device=dev)
    size = torch.zeros(b0.n, dtype=torch.int64, device=dev)
    ctx.decompressed_size_batch_device(torch.from_numpy(b0.src).to(dev), as_i64(b0.src_off), as_i64(b0.src_len),
                                       status, size)
    torch.cuda.synchronize(dev)
    assert int(status.max()) == 0
    b = T.Batch(streams, [int(x) for x in size.cpu().numpy()], dst_align=1)
    st, wr, dst = gpu_util.run_device(ctx, b)
    assert not st.any() and (wr == b.dst_cap).all()
    for k in (0, 7, 150, 299):
        assert b.dst_slice(dst, k).tobytes() == zlib.decompress(streams[k], -15)


def _chunk_plain(seed, n_words, total):
    rng = np.random.default_rng(seed)
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9))).astype(np.uint8)) for _ in range(n_words)]
    return b" ".join(words[int(i)] for i in rng.integers(0, n_words, total))


@pytest.mark.parametrize("piece,room", [(1 << 10, 1 << 20), (37_777, 1 << 20), (1 << 20, 1 << 21), (5_000, 70_000)])
def test_chunked_input_stream(ctx, oracle, piece, room):
    """sfb200_inflate_stream_*: a ~2.5 MB multi-block stream (dynamic, fixed and stored blocks, flush
    points) fed in pieces, the output taken in pieces — byte-identical to one call over the whole
    input (the oracle), whatever the piece size; dst too small for a block asks for more room."""
    plain = _chunk_plain(11, 400, 330_000) + bytes(np.random.default_rng(3).integers(0, 256, 150_000, dtype=np.uint8))
    plain += b"A" * 300_000 + _chunk_plain(12, 50, 60_000)
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = b""
    step = 180_000
    for k in range(0, len(plain), step):
        comp += co.compress(plain[k:k + step])
        comp += co.flush([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_NO_FLUSH][(k // step) % 3]) if k + step < len(plain) else b""
    comp += co.flush()
    want_st, want = oracle.decompress(comp, len(plain) + 16)[:2]
    assert want_st == 0 and bytes(want[:len(plain)]) == plain
    s = ctx.inflate_stream()
    try:
        out = bytearray()
        pos = 0
        fin = False
        guard = 0
        while not fin:
            chunk = comp[pos:pos + piece]
            pos += len(chunk)
            st, got, fin = s.feed(chunk, pos >= len(comp), room)
            out += got
            while not fin and st == 4 and not got:   # the next block does not fit: more room
                st, got, fin = s.feed(b"", pos >= len(comp), 1 << 21)
                out += got
            assert st == 0, (st, pos)
            guard += 1
            assert guard < 10_000
        assert bytes(out) == plain
        assert s.feed(b"", True, 16) == (0, b"", True)   # after the end: the final status again
    finally:
        s.close()


def test_chunked_input_stream_errors(ctx, oracle, golden):
    """Input that stops short, and a corrupt block, end the stream with the status (and the bytes)
    one call over the same input gives — once `last` says no more input will come."""
    plain = _chunk_plain(21, 300, 120_000)
    comp = zlib.compress(plain, 6)[2:-4]
    for name, data in (("cut", comp[:len(comp) * 2 // 3]), ("flip", comp[:40_000] + bytes([comp[40_000] ^ 0x5A]) + comp[40_001:])):
        want_st, want, want_wr = oracle.decompress(data, len(plain) + 64)[:3]
        s = ctx.inflate_stream()
        try:
            out = bytearray()
            fin = False
            pos = 0
            while not fin:
                chunk = data[pos:pos + 9_000]
                pos += len(chunk)
                st, got, fin = s.feed(chunk, pos >= len(data), 1 << 20)
                out += got
                if not fin:
                    assert st == 0
            assert st == want_st, (name, st, want_st)
            assert bytes(out) == bytes(want[:want_wr]), name
        finally:
            s.close()


def test_compress_round_trip_batch(ctx, oracle):
    """sfb200_compress_batch_device (§8 f4): a ragged batch — text, runs, random bytes, empty and tiny
    streams, odd alignments — compressed on the GPU, then decoded three ways: by this library's
    decompressor, by zlib and by the oracle (the reference's algorithm).  All give the input back."""
    rng = np.random.default_rng(8)
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9))).astype(np.uint8)) for _ in range(500)]
    plains = []
    for i in range(300):
        k = i % 6
        n = int(rng.integers(0, 70_000)) if i % 7 else int(rng.integers(0, 12))
        if k == 0:
            p = b" ".join(words[int(j)] for j in rng.integers(0, 500, n // 6 + 1))[:n]
        elif k == 1:
            p = bytes([int(rng.integers(0, 256))]) * n
        elif k == 2:
            p = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        elif k == 3:
            unit = bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8))
            p = (unit * (n // len(unit) + 1))[:n]
        elif k == 4:
            p = bytes(rng.integers(144, 256, n, dtype=np.uint8) if n else b"")
        else:
            p = T.text_like(n, 1000 + i) if n else b""
        plains.append(p)
    plains.append(_chunk_plain(5, 300, 200_000))   # > 65 535 bytes: the 16-bit position table wraps
    n = len(plains)
    lens = np.array([len(p) for p in plains], dtype=np.uint64)
    src_off = np.zeros(n, np.uint64)
    src_off[1:] = np.cumsum(lens + np.uint64(3))[:-1]          # odd gaps: every alignment
    src = np.zeros(int(src_off[-1] + lens[-1]) + 8, np.uint8)
    for i, p in enumerate(plains):
        src[int(src_off[i]):int(src_off[i]) + len(p)] = np.frombuffer(p, dtype=np.uint8)
    caps = lens + lens // np.uint64(8) + np.uint64(64)
    dst_off = np.zeros(n, np.uint64)
    dst_off[1:] = np.cumsum(caps + np.uint64(1))[:-1]
    total = int(dst_off[-1] + caps[-1])
    dev = torch.device("cuda", ctx.device)
    as_i64 = lambda a: torch.from_numpy(a.view(np.int64)).to(dev)
    d_src = torch.from_numpy(src).to(dev)
    d_dst = torch.full((total,), 0xA5, dtype=torch.uint8, device=dev)
    d_st = torch.full((n,), 0xEE, dtype=torch.uint8, device=dev)
    d_wr = torch.zeros(n, dtype=torch.int64, device=dev)
    ctx.compress_batch_device(d_src, as_i64(src_off), as_i64(lens), d_dst, as_i64(dst_off), as_i64(caps), d_st, d_wr)
    torch.cuda.synchronize()
    st = d_st.cpu().numpy()
    wr = d_wr.cpu().numpy().view(np.uint64)
    comp_all = d_dst.cpu().numpy()
    assert (st == 0).all()
    assert (wr <= caps).all()
    comps = [comp_all[int(dst_off[i]):int(dst_off[i] + wr[i])].tobytes() for i in range(n)]
    for i in range(n):
        assert zlib.decompress(comps[i], -15) == plains[i], i
        ost, out, owr, ub = oracle.decompress(comps[i], len(plains[i]) + 8)
        assert (ost, owr, ub) == (0, len(plains[i]), 0) and out[:owr] == plains[i], i
    # ... and back through this library's own decoder, as one batch
    b = T.Batch(comps, [len(p) for p in plains])
    dst_st, dst_wr, dst = gpu_util.run_device(ctx, b)
    assert (dst_st == 0).all()
    for i in range(n):
        assert int(dst_wr[i]) == len(plains[i]) and b.dst_slice(dst, i).tobytes()[:len(plains[i])] == plains[i], i
    ratio = float(lens.sum()) / float(wr.sum())
    assert ratio > 1.2
    # text of some length gets a dynamic-Huffman block (BTYPE 10), random bytes stored ones (00)
    kinds = [(comps[i][0] >> 1) & 3 for i in range(n)]
    assert all(kinds[i] == 2 for i in range(n) if i % 6 in (0, 5) and len(plains[i]) > 2000)
    assert all(kinds[i] == 0 for i in range(n) if i % 6 == 2 and len(plains[i]) > 2000)
    text_in = sum(len(plains[i]) for i in range(n) if i % 6 == 5)
    text_out = sum(len(comps[i]) for i in range(n) if i % 6 == 5)
    text_z = sum(len(zlib.compress(plains[i], 6)) for i in range(n) if i % 6 == 5)
    assert text_out < 1.25 * text_z, (text_in, text_out, text_z)   # within a quarter of zlib level 6 on text


def test_compress_single_stream_host_api_and_small_dst(ctx):
    data = _chunk_plain(9, 200, 20_000)
    st, comp = ctx.compress(data)
    assert st == 0 and zlib.decompress(comp, -15) == data and len(comp) < len(data) * 0.7
    rnd = bytes(np.random.default_rng(4).integers(0, 256, 5000, dtype=np.uint8))
    st, comp = ctx.compress(rnd)                       # incompressible: stored blocks, inside the bound
    assert st == 0 and len(comp) == 5005 and zlib.decompress(comp, -15) == rnd
    st, comp = ctx.compress(rnd, 4000)
    assert st == 4 and comp == b""
    st, comp = ctx.compress(b"")
    assert st == 0 and zlib.decompress(comp, -15) == b""
    st2, out, wr = ctx.decompress(ctx.compress(data)[1], len(data))
    assert (st2, wr) == (0, len(data)) and out == data
