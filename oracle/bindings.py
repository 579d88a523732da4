"""TEST INFRASTRUCTURE ONLY — ctypes loaders for the CPU oracle.

`load_oracle()`  -> oracle/_build/liboracle.so   (plain-C restatement, inflate_oracle.c)
`load_reference()` -> oracle/_ref/libstarflate_ref.so (UNMODIFIED reference + ref_shim.cpp),
                      or None when it was never built (it needs /root/reference at build time).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package starflate_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libstarflate_ref.so")

STATUS_NAMES = {
    0: "Success", 1: "Error", 2: "InvalidBlockHeader", 3: "NoCompressionLenMismatch",
    4: "DstTooSmall", 5: "SrcTooSmall", 6: "InvalidLitOrLen", 7: "InvalidDistance",
}


def build(with_reference: bool | None = None) -> None:
    """Compile the C restatement; and the reference too when /root/reference exists."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if with_reference is None:
        with_reference = os.path.isdir("/root/reference/src")
    if with_reference:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class _Result(C.Structure):
    _fields_ = [("status", C.c_uint8), ("ref_undefined", C.c_uint8),
                ("written", C.c_uint64), ("bits_consumed", C.c_uint64)]


_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


class Oracle:
    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(with_reference=False)
        self.lib = C.CDLL(path)
        L = self.lib
        L.sfo_decompress.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.POINTER(_Result)]
        L.sfo_decompress.restype = None
        L.sfo_decompress_batch.argtypes = [_u8p, _u64p, _u64p, _u8p, _u64p, _u64p, _u8p, _u64p,
                                           _u8p, C.c_uint64, C.c_int]
        L.sfo_decompress_batch.restype = None
        L.sfo_read_header.argtypes = [_u8p, C.c_size_t, C.c_uint8, C.POINTER(C.c_int)]
        L.sfo_copy_from_before.argtypes = [_u8p, C.c_size_t, C.c_uint16, C.c_uint16]
        L.sfo_canonical_codes.argtypes = [_u8p, C.c_size_t, _u64p, C.POINTER(C.c_uint16)]
        L.sfo_canonical_codes.restype = C.c_size_t
        L.sfo_decode_one.argtypes = [_u8p, C.c_size_t, _u8p, C.c_uint64, C.c_uint64,
                                     C.POINTER(C.c_uint16)]
        L.sfo_decode_one.restype = C.c_int
        L.sfo_fnv1a64.argtypes = [_u8p, C.c_size_t]
        L.sfo_fnv1a64.restype = C.c_uint64

    def decompress(self, src: bytes, dst_cap: int, fill: int = 0xA5):
        """-> (status, dst bytes over the full capacity, written, ref_undefined)"""
        s = np.frombuffer(src, dtype=np.uint8) if len(src) else np.zeros(1, np.uint8)
        d = np.full(max(dst_cap, 1), fill, dtype=np.uint8)
        r = _Result()
        self.lib.sfo_decompress(_ptr(s, _u8p), len(src), _ptr(d, _u8p), dst_cap, C.byref(r))
        return r.status, d[:dst_cap].tobytes(), r.written, r.ref_undefined

    def block_starts(self, src: bytes, dst_cap: int, max_starts: int = 1 << 16):
        """Bit positions of the block headers the front-to-back decode reads."""
        self.lib.sfo_block_starts.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.POINTER(_Result), _u64p,
                                              C.c_size_t]
        self.lib.sfo_block_starts.restype = C.c_size_t
        s = np.frombuffer(src, dtype=np.uint8) if len(src) else np.zeros(1, np.uint8)
        d = np.zeros(max(dst_cap, 1), np.uint8)
        r = _Result()
        out = np.zeros(max_starts, np.uint64)
        k = self.lib.sfo_block_starts(_ptr(s, _u8p), len(src), _ptr(d, _u8p), dst_cap, C.byref(r),
                                      _ptr(out, _u64p), max_starts)
        return [int(x) for x in out[:k]]

    def fnv1a64(self, data) -> int:
        a = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
        return int(self.lib.sfo_fnv1a64(_ptr(a, _u8p), len(data)))

    def decompress_batch(self, src, src_off, src_len, dst, dst_off, dst_cap, threads=1):
        n = len(src_off)
        status = np.zeros(n, np.uint8)
        written = np.zeros(n, np.uint64)
        ub = np.zeros(n, np.uint8)
        self.lib.sfo_decompress_batch(_ptr(src, _u8p), _ptr(src_off, _u64p), _ptr(src_len, _u64p),
                                      _ptr(dst, _u8p), _ptr(dst_off, _u64p), _ptr(dst_cap, _u64p),
                                      _ptr(status, _u8p), _ptr(written, _u64p), _ptr(ub, _u8p),
                                      n, threads)
        return status, written, ub

    def read_header(self, data: bytes, bit_size: int, bit_offset: int = 0):
        s = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
        out = (C.c_int * 5)()
        self.lib.sfo_read_header(_ptr(s, _u8p), bit_size, bit_offset, out)
        return list(out)

    def copy_from_before(self, buf: bytearray, dst_index: int, distance: int, n: int) -> bytes:
        a = np.frombuffer(bytes(buf), dtype=np.uint8).copy()
        self.lib.sfo_copy_from_before(_ptr(a, _u8p), dst_index, distance, n)
        return a.tobytes()

    def canonical_codes(self, lens):
        l = np.asarray(lens, dtype=np.uint8)
        codes = np.zeros(len(l), np.uint64)
        order = np.zeros(len(l), np.uint16)
        k = self.lib.sfo_canonical_codes(_ptr(l, _u8p), len(l), _ptr(codes, _u64p),
                                         order.ctypes.data_as(C.POINTER(C.c_uint16)))
        return codes, order[:k]

    def decode_one(self, lens, data: bytes, bit_pos: int, bit_end: int):
        l = np.asarray(lens, dtype=np.uint8)
        s = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
        sym = C.c_uint16(0)
        used = self.lib.sfo_decode_one(_ptr(l, _u8p), len(l), _ptr(s, _u8p), bit_pos, bit_end,
                                       C.byref(sym))
        return used, sym.value


class Reference:
    """The unmodified reference (oracle/_ref/libstarflate_ref.so)."""

    def __init__(self, path: str = REF_SO):
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_decompress.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t]
        L.ref_decompress.restype = C.c_int
        L.ref_read_header.argtypes = [_u8p, C.c_size_t, C.c_uint8, C.POINTER(C.c_int)]
        L.ref_copy_from_before.argtypes = [_u8p, C.c_size_t, C.c_size_t, C.c_uint16, C.c_uint16]
        L.ref_decompress_batch.argtypes = [_u8p, _u64p, _u64p, _u8p, _u64p, _u64p, _u8p,
                                           C.c_uint64, C.c_int]
        L.ref_decompress_batch.restype = None

    def decompress(self, src: bytes, dst_cap: int, fill: int = 0xA5):
        """Only call on inputs where the reference has defined behaviour."""
        s = np.frombuffer(src, dtype=np.uint8) if len(src) else np.zeros(1, np.uint8)
        d = np.full(max(dst_cap, 1), fill, dtype=np.uint8)
        st = self.lib.ref_decompress(_ptr(s, _u8p), len(src), _ptr(d, _u8p), dst_cap)
        return st, d[:dst_cap].tobytes()

    def decompress_batch(self, src, src_off, src_len, dst, dst_off, dst_cap, threads=1):
        n = len(src_off)
        status = np.zeros(n, np.uint8)
        self.lib.ref_decompress_batch(_ptr(src, _u8p), _ptr(src_off, _u64p), _ptr(src_len, _u64p),
                                      _ptr(dst, _u8p), _ptr(dst_off, _u64p), _ptr(dst_cap, _u64p),
                                      _ptr(status, _u8p), n, threads)
        return status

    def read_header(self, data: bytes, bit_size: int, bit_offset: int = 0):
        s = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
        out = (C.c_int * 5)()
        self.lib.ref_read_header(_ptr(s, _u8p), bit_size, bit_offset, out)
        return list(out)

    def copy_from_before(self, buf: bytearray, dst_index: int, distance: int, n: int) -> bytes:
        a = np.frombuffer(bytes(buf), dtype=np.uint8).copy()
        self.lib.ref_copy_from_before(_ptr(a, _u8p), len(a), dst_index, distance, n)
        return a.tobytes()


_oracle = None
_reference = None


def load_oracle() -> Oracle:
    global _oracle
    if _oracle is None:
        _oracle = Oracle()
    return _oracle


def load_reference() -> Reference | None:
    global _reference
    if _reference is None and os.path.exists(REF_SO):
        _reference = Reference()
    return _reference
