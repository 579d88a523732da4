"""The C++23 drop-in interface (starflate_b200/cpp: starflate::decompress, detail::read_header,
detail::copy_from_before, huffman::bit_span/table/decode_one) driven by the reference's own test
cases, restated in tests/cpp/decompress_test.cpp.  Host-only cases run in the GPU-less
container; the decompress() cases need cuda:0."""
import os
import subprocess

import pytest

from starflate_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "decompress_test")


@pytest.fixture(scope="module")
def test_binary():
    build.build_all()
    src = os.path.join(ROOT, "tests", "cpp", "decompress_test.cpp")
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    deps = [src, build.CXX_SO] + [os.path.join(d, f) for d, _, fs in os.walk(os.path.join(ROOT, "starflate_b200", "cpp")) for f in fs]
    if not os.path.exists(BIN) or any(os.path.getmtime(d) > os.path.getmtime(BIN) for d in deps):
        subprocess.check_call(["g++", "-std=c++23", "-O1", "-fno-exceptions", "-w", "-pthread",
                               "-I", os.path.join(ROOT, "starflate_b200", "cpp"), "-I", os.path.join(ROOT, "include"),
                               "-o", BIN, src, "-L", build.OUT, "-lstarflate", "-lstarflate_b200",
                               f"-Wl,-rpath,{build.OUT}"])
    return BIN


def test_reference_host_cases(test_binary):
    r = subprocess.run([test_binary, "host"], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr


def test_headers_expose_reference_names():
    hdr = open(os.path.join(ROOT, "starflate_b200", "cpp", "src", "decompress.hpp")).read()
    for name in ("enum class DecompressStatus", "InvalidBlockHeader", "NoCompressionLenMismatch",
                 "DstTooSmall", "SrcTooSmall", "InvalidLitOrLen", "InvalidDistance", "read_header",
                 "copy_from_before", "BlockHeader", "BlockType", "auto decompress("):
        assert name in hdr


@pytest.mark.gpu
def test_reference_decompress_cases_on_gpu(test_binary):
    r = subprocess.run([test_binary, "gpu", os.path.join(ROOT, "tests", "golden", "bases")],
                       capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
