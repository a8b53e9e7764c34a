"""In-tree build of liblwfb90.so (hand-written CUDA for sm_100a, C ABI of include/lwfb90.h)."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "liblwfb90.so")
_OBJ = os.path.join(_HERE, "_obj")

# -fmad=false: no implicit FMA contraction, so that the arithmetic is reproducible bit-for-bit by the
# host replay (oracle mode B); the elementary functions use explicit fma().
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC"]
# translation units: (source, extra flags); two units so that they compile in parallel.
# lwfb90_member.cu (thread-per-member fallback, not on the measured path) is built with NVVM -O1: at -O3 nvcc 12.9 has
# produced wrong results in this kernel twice (round 1: a select between the addresses of two local arrays; round 2:
# isoBEA2016 member 22 goes NaN in its first step while -O1, -O0, the host build of the same text and the lane kernel agree
# with the oracle bit for bit, tools/legacy_diag.py).  Both builds are checked against the oracle by the GPU tests.
UNITS = [("lwfb90.cu", []), ("lwfb90_member.cu", ["-Xcicc", "-O1"]), ("lwfb90_multi.cu", []),
         ("lwfb90_wide.cu", ["-DLWFB_CB_SCRATCH=1"])]


def _deps():
    return (glob.glob(os.path.join(_CSRC, "*.cu")) + glob.glob(os.path.join(_CSRC, "*.cuh")) + glob.glob(os.path.join(_CSRC, "*.h"))
            + glob.glob(os.path.join(os.path.dirname(_HERE), "include", "*.h")) + [os.path.abspath(__file__)])


def nvcc_path() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    return (not os.path.exists(OUT)) or any(os.path.getmtime(OUT) < os.path.getmtime(d) for d in _deps())


def build_library(force: bool = False, verbose: bool = False, extra=()) -> str:
    if force or stale():
        os.makedirs(_OBJ, exist_ok=True)
        nvcc = nvcc_path()
        procs, objs = [], []
        for src, flags in UNITS:
            obj = os.path.join(_OBJ, src.replace(".cu", ".o"))
            objs.append(obj)
            cmd = [nvcc] + NVCC_FLAGS + list(flags) + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
            procs.append((cmd, subprocess.Popen(cmd, cwd=_CSRC)))
        for cmd, p in procs:
            if p.wait() != 0:
                raise subprocess.CalledProcessError(p.returncode, cmd)
        tmp = OUT + ".tmp"      # link beside the target and rename: a reader never sees a half-written library
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + objs, cwd=_CSRC)
        os.replace(tmp, OUT)
    return OUT
