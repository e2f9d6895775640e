"""Builds photonbend_b200/libpbremap.so (the CUDA kernels + C ABI) in-tree with nvcc.

    python -m photonbend_b200.build [--force]

sm_100a only.  -fmad=false / -ffp-contract=off: the reference never fuses a multiply with an
add, and the truncated source index has to match its float64 arithmetic bit for bit.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_PATH = os.path.join(PKG, "libpbremap.so")
SOURCES = [os.path.join(CSRC, "pb_remap.cu")]
HEADERS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [
    os.path.join(REPO, "include", "pb_remap.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
    "-shared",
]


IO_LIB_PATH = os.path.join(PKG, "libpbio.so")
IO_SOURCE = os.path.join(CSRC, "pb_io.cpp")
IO_HEADER = os.path.join(REPO, "include", "pb_io.h")


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libpbremap.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False, out: str = LIB_PATH) -> str:
    """out != LIB_PATH: an experiment build next to the product library (loaded with PB_REMAP_LIB)."""
    if out == LIB_PATH and not force and not is_stale():
        return LIB_PATH
    extra = os.environ.get("PB_NVCC_EXTRA", "").split()  # e.g. -DPB_EXPERIMENTS for timing experiments
    cmd = [find_nvcc(), *NVCC_FLAGS, *extra, "-I", os.path.join(REPO, "include"), "-I", CSRC,
           "-o", out, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return out


def io_is_stale() -> bool:
    if not os.path.exists(IO_LIB_PATH):
        return True
    built = os.path.getmtime(IO_LIB_PATH)
    return any(os.path.getmtime(p) > built for p in (IO_SOURCE, IO_HEADER))


def build_io(force: bool = False) -> str:
    """libpbio.so: the nvJPEG bridge (host C++, no kernels of its own)."""
    if not force and not io_is_stale():
        return IO_LIB_PATH
    cuda = os.path.dirname(os.path.dirname(find_nvcc()))
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    cmd = [cxx, "-O2", "-fPIC", "-shared", "-std=c++17", "-I", os.path.join(cuda, "include"),
           "-I", os.path.join(REPO, "include"), IO_SOURCE, "-L", os.path.join(cuda, "lib64"),
           "-lnvjpeg", "-lcudart", "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", IO_LIB_PATH]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building libpbio.so failed:\n" + proc.stdout + proc.stderr)
    return IO_LIB_PATH


if __name__ == "__main__":
    print(build_io(force="--force" in sys.argv))
    out = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")), LIB_PATH)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=out))
