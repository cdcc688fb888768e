"""Builds liborbx.so (the C-ABI CUDA library, include/orbx.h) in-tree with nvcc for sm_100a.

    python -m orbslam2_with_quadrics_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "liborbx.so")
SOURCES = ["orbx_kernels.cu", "orbx_api.cu"]
DEPS = SOURCES + ["orbx_plan.h", "orbx_kernels.h", "orb_pattern_31.inc", os.path.join("..", "..", "include", "orbx.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "--fmad=false",                       # canonical rule B-2: nothing is contracted into FMAs
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", "-shared", "-cudart", "static",
]


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


LIB_BC = os.path.join(PKG, "liborbx_boundscheck.so")     # -DORBX_BOUNDS_CHECK: device-side traps on every tile / queue / map index


def stale(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_library(force: bool = False, verbose: bool = False, bounds_check: bool = False) -> str:
    lib = LIB_BC if bounds_check else LIB
    if not force and not stale(lib):
        return lib
    cmd = [nvcc()] + NVCC_FLAGS + (["-DORBX_BOUNDS_CHECK"] if bounds_check else []) + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, s) for s in SOURCES] + ["-o", lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return lib


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_library(force="--force" in sys.argv, bounds_check=True))
