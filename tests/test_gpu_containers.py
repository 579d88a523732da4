"""GPU: zlib (RFC 1950) / gzip (RFC 1952) containers through the C ABI.  The
reference is raw-DEFLATE only, so the checker here is zlib itself (Python's zlib / gzip modules):
every container they write must decode to the same bytes with status Success, and every way of
breaking one must be reported as the RFCs name it."""
import gzip
import io
import struct
import zlib

import numpy as np
import pytest
import torch

from tests import deflate_tools as T
from tests import gpu_util

pytestmark = pytest.mark.gpu

BAD, CHECKSUM, SIZE = 8, 9, 10


def _gzip(data, level=6, name=None, extra=None, comment=None, hcrc=False):
    """A gzip member with any combination of the optional header fields (RFC 1952 §2.3)."""
    flg = (4 if extra is not None else 0) | (8 if name is not None else 0) | (16 if comment is not None else 0) | \
          (2 if hcrc else 0)
    h = bytearray(b"\x1f\x8b\x08" + bytes([flg]) + struct.pack("<I", 1234567) + b"\x00\x03")
    if extra is not None:
        h += struct.pack("<H", len(extra)) + extra
    if name is not None:
        h += name + b"\0"
    if comment is not None:
        h += comment + b"\0"
    if hcrc:
        h += struct.pack("<H", zlib.crc32(bytes(h)) & 0xffff)
    return bytes(h) + T.raw_deflate(data, level) + struct.pack("<II", zlib.crc32(data), len(data) & 0xffffffff)


def _run(ctx, container, streams, caps):
    dev = torch.device("cuda", ctx.device)
    b = T.Batch(streams, caps, dst_align=1, src_align=1)
    as_i64 = lambda a: torch.from_numpy(a.view(np.int64)).to(dev)
    dst = torch.full((b.dst_total,), 0xA5, dtype=torch.uint8, device=dev)
    status = torch.full((b.n,), 0xEE, dtype=torch.uint8, device=dev)
    written = torch.full((b.n,), -1, dtype=torch.int64, device=dev)
    ctx.decompress_container_batch_device(container, torch.from_numpy(b.src).to(dev), as_i64(b.src_off),
                                          as_i64(b.src_len), dst, as_i64(b.dst_off), as_i64(b.dst_cap), status, written)
    torch.cuda.synchronize(dev)
    return b, status.cpu().numpy(), written.cpu().numpy(), dst.cpu().numpy()


def _plains():
    out = [b"", b"a", b"hello, hello, hello", bytes(range(256)) * 3]
    for i, (kind, size) in enumerate((("dynamic", 5000), ("fixed", 777), ("stored", 3000), ("repetitive", 70000),
                                      ("multiblock", 40000), ("dynamic", 300001))):
        out.append(T.make_stream(kind, size, 600 + i)[0])
    return out


def test_zlib_and_gzip_written_by_zlib_decode_and_verify(ctx):
    plains = _plains()
    zs = [zlib.compress(p, lvl) for p in plains for lvl in (1, 6, 9)]
    gs = [gzip.compress(p, compresslevel=lvl, mtime=0) for p in plains for lvl in (1, 9)]
    buf = io.BytesIO()
    with gzip.GzipFile(filename="some file.txt", mode="wb", fileobj=buf, mtime=99) as f:  # FNAME set by the library
        f.write(plains[5])
    gs.append(buf.getvalue())
    gs += [_gzip(plains[4], name=b"n", comment=b"a comment", extra=b"\x01\x02\x03\x04\x05", hcrc=True),
           _gzip(plains[6], extra=b"", hcrc=True), _gzip(plains[2], comment=b""), _gzip(plains[0], hcrc=True)]
    for container, streams in ((ctx.ZLIB, zs), (ctx.GZIP, gs), (ctx.AUTO, zs + gs)):
        want = [zlib.decompress(s, 15 + 32) for s in streams]  # zlib's own auto-detecting decoder
        b, st, wr, dst = _run(ctx, container, streams, [len(w) + 5 for w in want])
        assert not st.any(), (container, np.nonzero(st)[0][:5], st[st != 0][:5])
        for k, w in enumerate(want):
            got = b.dst_slice(dst, k).tobytes()
            assert int(wr[k]) == len(w) and got[: len(w)] == w and got[len(w):] == b"\xa5" * 5, (container, k)


def test_auto_takes_raw_deflate_for_raw(ctx):
    plains = _plains()
    raws = [T.raw_deflate(p, 6) for p in plains]
    raws = [r for r in raws if not (r[:2] == b"\x1f\x8b" or ((r[0] & 15) == 8 and (r[0] >> 4) <= 7 and
                                                             ((r[0] << 8) | r[1]) % 31 == 0))]
    assert len(raws) >= 8
    b, st, wr, dst = _run(ctx, ctx.AUTO, raws, [len(zlib.decompress(r, -15)) for r in raws])
    assert not st.any()
    for k, r in enumerate(raws):
        assert b.dst_slice(dst, k).tobytes() == zlib.decompress(r, -15)


def test_broken_containers_are_reported(ctx, oracle):
    p = T.make_stream("dynamic", 20000, 5)[0]
    z, g = zlib.compress(p, 6), _gzip(p, name=b"x")
    flip = lambda s, i, m=0x01: s[:i] + bytes([s[i] ^ m]) + s[i + 1:]
    zl = [
        (z, len(p), 0), (flip(z, len(z) - 1), len(p), CHECKSUM), (flip(z, len(z) - 4, 0x80), len(p), CHECKSUM),
        (flip(z, 1), len(p), BAD),                      # FCHECK no longer fits
        (bytes([0x79, 0x9c]) + z[2:], len(p), BAD),     # CM = 9
        (bytes([0x78, 0xbb]) + z[2:], len(p), BAD),     # FDICT (0x78bb % 31 == 0)
        (z[:5], len(p), BAD), (b"", len(p), BAD),
        (z, len(p) - 1, 4),                             # dst too small: the DEFLATE status stands
        (z[:2] + z[2:-4][:1000] + z[-4:], len(p), None),  # payload cut: whatever the raw decoder says
    ]
    gz = [
        (g, len(p), 0), (flip(g, len(g) - 8), len(p), CHECKSUM), (flip(g, len(g) - 1, 0x40), len(p), SIZE),
        (flip(g, 0), len(p), BAD), (flip(g, 2), len(p), BAD), (flip(g, 3, 0x80), len(p), BAD),  # magic, CM, reserved flag
        (g[:12], len(p), BAD),
        (_gzip(p, hcrc=True)[:10] + b"\x00\x00" + _gzip(p, hcrc=True)[12:], len(p), BAD),      # wrong header CRC16
        (b"\x1f\x8b\x08\x08" + b"\0" * 6 + b"name-without-end" * 3, 10, BAD),                   # FNAME runs into the trailer
        (g, 100, 4),
    ]
    for container, cases in ((ctx.ZLIB, zl), (ctx.GZIP, gz)):
        b, st, wr, dst = _run(ctx, container, [c[0] for c in cases], [c[1] for c in cases])
        for k, (s, cap, want) in enumerate(cases):
            if want is None:
                want = oracle.decompress(s[2:-4], cap)[0]
                assert want != 0
            assert int(st[k]) == want, (container, k, int(st[k]), want)
            if want == BAD:
                assert int(wr[k]) == 0 and (b.dst_slice(dst, k) == 0xA5).all()
            if want in (0, CHECKSUM, SIZE):
                assert b.dst_slice(dst, k).tobytes() == p  # the bytes are there either way


def test_single_container_host_api_and_large_stream(ctx):
    """sfb200_decompress_container (host buffers), incl. a 12 MiB gzip member that goes down the
    single-stream route (CRC-32 over 12 MiB by one warp per stream)."""
    p = T.big_text(12 << 20, 3)
    for container, blob in ((ctx.GZIP, gzip.compress(p, compresslevel=6, mtime=0)), (ctx.ZLIB, zlib.compress(p, 6)),
                            (ctx.AUTO, zlib.compress(p[:70000], 9))):
        want = zlib.decompress(blob, 47)
        st, dst, wr = ctx.decompress_container(container, blob, len(want) + 3, fill=0x5A)
        assert (st, wr) == (0, len(want)) and dst[: len(want)] == want and dst[len(want):] == b"\x5a" * 3
    st, dst, wr = ctx.decompress_container(ctx.GZIP, gzip.compress(p[:5000])[:-3] + b"abc", 5000)
    assert st in (CHECKSUM, SIZE)
