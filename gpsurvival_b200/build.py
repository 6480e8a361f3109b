"""Build libgpsurv.so (the CUDA C-ABI library) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting
``gpsurvival_b200/libgpsurv.so`` travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgpsurv.so")
SOURCES = [os.path.join(CSRC, "gpsurv.cu")]


def deps():
    """Everything the library is built from: every file under csrc/ (gpsurv.cu includes all the .cuh files) and
    the public header — a stale .so must never travel to the GPU box."""
    out = [os.path.join(os.path.dirname(HERE), "include", "gpsurv.h"), os.path.abspath(__file__)]
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
    return out

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in deps() if os.path.exists(f))


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
