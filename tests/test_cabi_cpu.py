"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/*.h declares,
contains the kernels as sm_100a SASS, and refuses to work without a device (no CPU fallback)."""
import os
import re
import subprocess

import pytest

import starflate_b200 as S
from starflate_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build_cabi()


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "starflate_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sfb200_[a-z_0-9]+)\s*\(", hdr)))


def test_exports_every_declared_symbol(lib_path):
    syms = _declared_symbols()
    assert "sfb200_decompress_batch_device" in syms and len(syms) >= 9
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True,
                         check=True).stdout
    exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert not [s for s in syms if s not in exported]
    lib = S.load_library()
    for s in syms:
        assert getattr(lib, s) is not None
    assert lib.sfb200_abi_version() == 6


def test_no_cpu_fallback_without_device(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(S.StarflateError, match="NO_DEVICE"):
        S.Context(0)


def test_product_does_not_reference_oracle():
    """The product package must never import / link the oracle or the CPU emulation."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "starflate_b200")):
        if "_build" in d or "__pycache__" in d:
            continue
        for f in files:
            text = open(os.path.join(d, f), errors="ignore").read()
            if re.search(r"oracle|liboracle|inflate_oracle|emu_bindings", text):
                bad.append(os.path.join(d, f))
    assert not bad, bad
    out = subprocess.run(["ldd", build.CABI_SO], capture_output=True, text=True).stdout
    assert "oracle" not in out and "emu" not in out


def test_kernels_are_sm100a_sass(lib_path):
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True).stdout
    assert re.search(r"Function : \S*huff_lanes_kernel", sass)
    assert re.search(r"Function : \S*lz_window_kernel", sass)
    for name in ("huff_stream_kernel", "find_candidates_kernel", "verify_candidates_kernel", "chain_kernel",
                 "lz_jump_init_kernel", "lz_jump_round_kernel"):  # the single-stream route
        assert re.search(r"Function : \S*" + name, sass), name
    assert "LDS" in sass and "LDG" in sass and "STG" in sass
