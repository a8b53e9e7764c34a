"""ctypes wrapper of the lock-step warp emulator (tests/simt/simt_emul.cpp).  TEST INFRASTRUCTURE ONLY.

Runs the *kernel text* of lwfbrook90.jl_b200/csrc on the host with SIMT semantics so that the non-GPU
test suite can compare it bit-for-bit with the CPU oracle.  Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_SO = os.path.join(_HERE, "_build", "libsimt_emul.so")
_SO_WIDE = os.path.join(_HERE, "_build", "libsimt_emul_wide.so")     # the warp-per-member kernel (N > 32): own build, see simt_emul.cpp
_lib = None
_lib_wide = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)

LANES, MEMBER, WIDE = 0, 1, 2


def build(force: bool = False, wide: bool = False) -> str:
    _SO = _SO_WIDE if wide else globals()["_SO"]
    deps = [os.path.join(_HERE, "simt_emul.cpp")] + glob.glob(os.path.join(_ROOT, "lwfbrook90.jl_b200", "csrc", "*.cuh")) \
        + glob.glob(os.path.join(_ROOT, "lwfbrook90.jl_b200", "csrc", "*.h")) + glob.glob(os.path.join(_ROOT, "include", "*.h"))
    if force or not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(f) for f in deps):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        fma = ["-mfma"] if "fma" in open("/proc/cpuinfo").read().split() else []
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fno-fast-math", "-ffp-contract=off", *fma, "-fPIC", "-shared",
                               *(["-DLWFB_EMUL_WIDE=1"] if wide else []), *os.environ.get("LWFB_EMUL_CFLAGS", "").split(), "-o", _SO, os.path.join(_HERE, "simt_emul.cpp"), "-lpthread"])
    return _SO


def _load(path):
    L = C.CDLL(path)
    L.lwfb_emul_run.restype = C.c_int
    L.lwfb_emul_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int64, _ip,
                                _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int,
                                _dp, _ip, _lp, _ip, C.c_int, C.c_int, C.c_int64, C.c_int]
    return L


_current_wide = False     # which build the helper calls (set_observations, set_ntfli0, eval_rhs ...) address


def lib(wide=None):
    global _lib, _lib_wide
    if wide is None:
        wide = _current_wide
    if wide:
        if _lib_wide is None:
            _lib_wide = _load(build(wide=True))
        return _lib_wide
    if _lib is None:
        _lib = _load(build())
    return _lib


def use_wide(on: bool):
    """Route the helper calls and run() to the build that holds the warp-per-member kernel (N > 32)."""
    global _current_wide
    _current_wide = bool(on)


def _p(a, t=_dp):
    return a.ctypes.data_as(t)


def set_observations(n_days, obs_day=None, obs_layer=None, obs_theta=None, obs_d18o=None):
    """The emulator's lwfb_set_observations: the following runs accumulate the score sums (score_sums())."""
    L = lib()
    L.lwfb_emul_set_obs.argtypes = [C.c_int, C.c_int, _ip, C.c_int, _ip, _dp, _dp]
    L.lwfb_emul_set_obs.restype = None
    if obs_day is None:
        z = np.zeros(1, dtype=np.int32)
        L.lwfb_emul_set_obs(int(n_days), 0, _p(z, _ip), 0, _p(z, _ip), None, None)
        return
    od = np.ascontiguousarray(obs_day, dtype=np.int32)
    ol = np.ascontiguousarray(obs_layer, dtype=np.int32)
    th = None if obs_theta is None else np.ascontiguousarray(obs_theta, dtype=np.float64)
    d18 = None if obs_d18o is None else np.ascontiguousarray(obs_d18o, dtype=np.float64)
    L.lwfb_emul_set_obs(int(n_days), len(od), _p(od, _ip), len(ol), _p(ol, _ip), None if th is None else _p(th),
                        None if d18 is None else _p(d18))


def set_ntfli0(ntfli0=None):
    """du_NTFLI cache seen by the first ITER of the following runs ([M, N]); None: fresh problem."""
    L = lib()
    L.lwfb_emul_set_ntfli0.argtypes = [_dp, C.c_int64]; L.lwfb_emul_set_ntfli0.restype = None
    if ntfli0 is None:
        L.lwfb_emul_set_ntfli0(None, 0)
    else:
        a = np.ascontiguousarray(ntfli0, dtype=np.float64)
        L.lwfb_emul_set_ntfli0(_p(a), a.size)


def get_ntfli(M, N):
    L = lib()
    L.lwfb_emul_get_ntfli.argtypes = [_dp, C.c_int64]; L.lwfb_emul_get_ntfli.restype = None
    out = np.zeros((M, N))
    L.lwfb_emul_get_ntfli(_p(out), out.size)
    return out


def eval_rhs(N, iso, forcing, site_of_member, scal, lay, u0, t, swati, gwat, frozen, ntfli_prev, warps_per_block=1, kernel=LANES):
    """One evaluation of f(u, t) by the LANE kernel's own stage code (the probe behind lwfb_eval_rhs for 3 <= N <= 32):
    du [M, N+9]."""
    L = lib()
    L.lwfb_emul_set_probe.argtypes = [_dp, _dp, C.c_double, _dp]; L.lwfb_emul_set_probe.restype = None
    u0p = np.ascontiguousarray(u0, dtype=np.float64).copy()
    M = u0p.shape[0]
    k = 3 if iso else 1
    u0p[:, :N] = swati
    u0p[:, k * N] = gwat
    fr = np.ascontiguousarray(frozen, dtype=np.float64); nt = np.ascontiguousarray(ntfli_prev, dtype=np.float64)
    du = np.full((M, N + 9), np.nan)
    L.lwfb_emul_set_probe(_p(fr), _p(nt), float(t), _p(du))
    try:
        run(kernel, N, iso, 0, 1, forcing, site_of_member, scal, lay, u0p, seg_len=1 if kernel == LANES else 0, warps_per_block=warps_per_block)
    finally:
        L.lwfb_emul_set_probe(None, None, 0.0, None)
    return du


def score_sums(M):
    """[M][2][7] running sums (n, se, so, ss, soo, sss, sos for theta and delta18O) of the last run."""
    L = lib()
    L.lwfb_emul_get_sums.argtypes = [_dp, C.c_int64]
    L.lwfb_emul_get_sums.restype = None
    out = np.zeros((M, 2, 7))
    L.lwfb_emul_get_sums(_p(out), out.size)
    return out


def run(kernel, N, iso, t0_day, n_days, forcing, site_of_member, scal, lay, u0, *, reltol=1e-5, abstol=1e-6,
        dt_init=1e-3, dtmax=60 / 60 / 24, maxiters=1_000_000, n_threads=8, perm=None, seg_len=0, resort=False,
        warps_per_block=1):
    """Same shapes and return values as oracle.run.  perm: optional placement (slot -> member) of the lane kernel;
    seg_len > 0: work-quantum execution (a launch ends after seg_len passes of the warp loop), resort: regroup the
    unfinished members by progress between the launches (resort=2: additionally with the stop flag that ends a launch
    when the first block has used up its quantum and no block is left to take its place).  warps_per_block: size of the emulated thread block (the
    block-uniform votes and stage barriers of the kernel span it; 8 = what the GPU runs)."""
    forcing = np.ascontiguousarray(forcing, dtype=np.float64)
    if forcing.ndim == 2:
        forcing = forcing[None]
    n_sites, n_rows, _ = forcing.shape
    scal = np.ascontiguousarray(scal, dtype=np.float64)
    lay = np.ascontiguousarray(lay, dtype=np.float64)
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    M = scal.shape[0]
    L = 9 * N + 51 if iso else 5 * N + 39
    assert u0.shape == (M, L) and lay.shape[0] == M and lay.shape[2] == N
    som = np.ascontiguousarray(site_of_member, dtype=np.int32)
    out = np.empty((M, n_days + 1, L))
    status = np.zeros(M, dtype=np.int32)
    counters = np.zeros((M, 3), dtype=np.int64)
    assert (kernel == WIDE) == _current_wide, "use_wide(True) selects the build that holds the warp-per-member kernel"
    rc = lib().lwfb_emul_run(int(kernel), N, int(iso), int(t0_day), int(n_days), n_sites, n_rows, _p(forcing), M, _p(som, _ip),
                             _p(scal), _p(lay), _p(u0), reltol, abstol, dt_init, dtmax, maxiters, int(n_threads), _p(out),
                             _p(status, _ip), _p(counters, _lp),
                             None if perm is None else _p(np.ascontiguousarray(perm, dtype=np.int32), _ip),
                             int(seg_len), int(resort), 0 if perm is None else int(len(perm)),
                             int(warps_per_block))
    if rc != 0:
        raise RuntimeError(f"lwfb_emul_run failed ({rc})")
    return out, status, counters
