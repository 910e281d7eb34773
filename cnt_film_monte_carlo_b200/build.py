"""Build the engine in-tree: libcntmc.so (nvcc, sm_100a) and the C++ driver cntmc_main (g++, links the library).

    python -m cnt_film_monte_carlo_b200.build [--force] [--verbose]
    python -m cnt_film_monte_carlo_b200.build --segments      # diagnostics twin libcntmc_seg.so (tools/segments.py)

-fmad=false and -ffp-contract=off are part of the arithmetic contract (bit parity with the reference needs unfused
multiply-adds; the path is not FLOP-bound, so this costs nothing measurable).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
CPP = os.path.join(HERE, "cpp")
LIB = os.path.join(HERE, "libcntmc.so")
DRIVER = os.path.join(HERE, "cntmc_main")
SOURCES = ["cntmc_api.cu", "cntmc_multi.cu", "cntmc_davoody.cu", "host_setup.cpp"]
HEADERS = ["hop_core.h", "fast_log.h", "log_table.inc", "csr_core.h", "kernels.cuh", "host_setup.h", "json_min.h",
           "herm_eig.h", "davoody_tube.h", "davoody_transfer.h", "davoody_kernels.cuh"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unknown-pragmas", "-shared", "-ccbin", "/usr/bin/g++",
]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "..", "include", "cntmc.h")]
    return _stale(LIB, deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        cmd = [NVCC, *FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES], "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    build_driver(force)
    return LIB


def build_segments() -> str:
    """Diagnostics build with per-segment cycle counters in the hop loop (tools/segments.py): libcntmc_seg.so."""
    out = os.path.join(HERE, "libcntmc_seg.so")
    subprocess.check_call([NVCC, *FLAGS, "-DCNTMC_PROFILE_SEGMENTS", "-o", out, *[os.path.join(CSRC, s) for s in SOURCES], "-ldl"])
    return out


def build_variant(name: str, *defines: str) -> str:
    """A/B builds for kernel experiments (tools/gpu_*.sh select them with CNTMC_LIB=...): libcntmc_<name>.so."""
    out = os.path.join(HERE, "libcntmc_%s.so" % name)
    subprocess.check_call([NVCC, *FLAGS, *["-D" + d for d in defines], "-o", out, *[os.path.join(CSRC, s) for s in SOURCES], "-ldl"])
    return out


def build_driver(force: bool = False) -> str:
    """cpp/main.cpp + cpp/monte_carlo.hpp: the reference-shaped C++ host side over the C ABI."""
    deps = [os.path.join(CPP, "main.cpp"), os.path.join(CPP, "monte_carlo.hpp"), os.path.join(CSRC, "json_min.h"),
            os.path.join(HERE, "..", "include", "cntmc.h"), LIB]
    if force or _stale(DRIVER, deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-o", DRIVER, os.path.join(CPP, "main.cpp"),
                               "-L" + HERE, "-lcntmc", "-Wl,-rpath,$ORIGIN", "-lstdc++fs"])
    return DRIVER


if __name__ == "__main__":
    if "--segments" in sys.argv:
        print(build_segments())
        sys.exit(0)
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
    print(DRIVER)
