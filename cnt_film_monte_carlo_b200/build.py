"""Build libcntmc.so (the C-ABI engine) in-tree with nvcc for sm_100a.

    python -m cnt_film_monte_carlo_b200.build [--force] [--verbose]

-fmad=false and -ffp-contract=off are part of the arithmetic contract (bit parity with the reference needs unfused
multiply-adds; the path is gather-bound, so this costs nothing measurable).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcntmc.so")
SOURCES = ["cntmc_api.cu", "host_setup.cpp"]
HEADERS = ["hop_core.h", "csr_core.h", "kernels.cuh", "host_setup.h", "json_min.h"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-shared", "-ccbin", "/usr/bin/g++",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "..", "include", "cntmc.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [NVCC, *FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
