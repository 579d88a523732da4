"""Build the native pieces in-tree (sm_100a only):

  starflate_b200/_build/libstarflate_b200.so   CUDA kernels + C ABI   (nvcc, C++20)
  starflate_b200/_build/libstarflate.so        C++23 mirror of the reference interface (g++)

nvcc cross-compiles without a GPU, so this also runs in the GPU-less authoring container.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build")
CABI_SO = os.path.join(OUT, "libstarflate_b200.so")
CXX_SO = os.path.join(OUT, "libstarflate.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++20",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _gxx() -> str:
    for cand in ("/usr/bin/g++", shutil.which("g++")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("g++ not found")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cabi(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    csrc = os.path.join(HERE, "csrc")
    sources = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc))]
    sources.append(os.path.join(ROOT, "include", "starflate_b200.h"))
    if force or _stale(CABI_SO, sources):
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", CABI_SO, os.path.join(csrc, "capi.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    return CABI_SO


def build_cxx(force: bool = False) -> str | None:
    """C++23 host mirror (starflate::decompress & friends) on top of the C ABI."""
    cpp = os.path.join(HERE, "cpp")
    src = os.path.join(cpp, "src", "decompress.cpp")
    if not os.path.exists(src):
        return None
    sources = []
    for d, _, files in os.walk(cpp):
        sources += [os.path.join(d, f) for f in files]
    if force or _stale(CXX_SO, sources + [CABI_SO]):
        subprocess.check_call([_gxx(), "-std=c++23", "-O2", "-fPIC", "-shared", "-fno-exceptions",
                               "-I", cpp, "-I", os.path.join(ROOT, "include"), "-o", CXX_SO, src,
                               "-L", OUT, "-lstarflate_b200", "-Wl,-rpath,$ORIGIN"])
    return CXX_SO


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cabi(force, verbose)
    build_cxx(force)


if __name__ == "__main__":
    build_all(force=True, verbose=True)
