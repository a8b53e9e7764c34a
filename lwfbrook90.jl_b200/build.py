"""In-tree build of liblwfb90.so (hand-written CUDA for sm_100a, C ABI of include/lwfb90.h)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "lwfb90.cu")
OUT = os.path.join(_HERE, "liblwfb90.so")
DEPS = [SRC, os.path.join(_HERE, "csrc", "lwfb90_device.cuh"), os.path.join(_HERE, "csrc", "lwfb90_math.cuh"),
        os.path.join(os.path.dirname(_HERE), "include", "lwfb90.h"),
        os.path.join(os.path.dirname(_HERE), "include", "lwfb90_elem.h")]

# -fmad=false: no implicit FMA contraction, so that the arithmetic is reproducible bit-for-bit by the
# host replay (oracle mode B); the elementary functions use explicit fma().
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-shared", "-Xcompiler", "-fPIC"]


def nvcc_path() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    return (not os.path.exists(OUT)) or any(os.path.getmtime(OUT) < os.path.getmtime(d) for d in DEPS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
        subprocess.check_call(cmd, cwd=os.path.dirname(SRC))
    return OUT
