"""ctypes wrapper of the CPU oracle (oracle/lwfb90_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblwfb90_oracle.so")
_SO_ELEM = os.path.join(_HERE, "_build", "liblwfb90_oracle_elem.so")
_lib = None
_libs = {}
_mode = "libm"


def use(mode: str):
    """Select the oracle build: "libm" (mode A, independent, pinned to goldens) or "elem"
    (mode B, elementary functions of include/lwfb90_elem.h = bit-exact replay of the GPU)."""
    global _mode, _lib
    assert mode in ("libm", "elem")
    _mode = mode
    _lib = _libs.get(mode)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, "lwfb90_oracle.c"), os.path.join(os.path.dirname(_HERE), "include", "lwfb90.h"),
           os.path.join(os.path.dirname(_HERE), "include", "lwfb90_elem.h"),
           os.path.join(os.path.dirname(_HERE), "include", "lwfb90_setup.h"), os.path.join(_HERE, "lwfb90_setup_twin.c")]
    stale = any(not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(f) for f in src)
                for so in (_SO, _SO_ELEM))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO if _mode == "libm" else _SO_ELEM)
        _libs[_mode] = _lib
        _lib.lwfb_oracle_run.restype = C.c_int
        _lib.lwfb_oracle_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp,
                                         C.c_int64, _ip, _dp, _dp, _dp, C.c_double, C.c_double,
                                         C.c_double, C.c_double, C.c_int64, C.c_int, _dp, _ip, _lp]
        _lib.lwfb_oracle_run_member.restype = C.c_int
        _lib.lwfb_oracle_run_member.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp,
                                                _dp, C.c_double, C.c_double, C.c_double, C.c_double,
                                                C.c_int64, _dp, _dp, _lp, _dp]
        _lib.lwfb_oracle_eval_rhs.restype = None
        _lib.lwfb_oracle_eval_rhs.argtypes = [C.c_int, _dp, _dp, C.c_double, _dp, C.c_double, _dp, _dp, _dp]
        _lib.lwfb_oracle_kkmean.restype = C.c_double
        _lib.lwfb_oracle_kkmean.argtypes = [C.c_double] * 4
        _lib.lwfb_oracle_pm.restype = C.c_double
        _lib.lwfb_oracle_pm.argtypes = [C.c_double] * 5
        _lib.lwfb_oracle_esat.restype = None
        _lib.lwfb_oracle_esat.argtypes = [C.c_double, _dp, _dp]
        _lib.lwfb_oracle_fastpow.restype = C.c_double
        _lib.lwfb_oracle_fastpow.argtypes = [C.c_double, C.c_double]
        _lib.lwfb_oracle_elem.restype = C.c_double
        _lib.lwfb_oracle_elem.argtypes = [C.c_int, C.c_double, C.c_double]
    return _lib


def _p(a, t=_dp):
    return a.ctypes.data_as(t)


def run(N, iso, t0_day, n_days, forcing, site_of_member, scal, lay, u0, *, reltol=1e-5, abstol=1e-6,
        dt_init=1e-3, dtmax=60 / 60 / 24, maxiters=1_000_000, n_threads=1, record=True):
    """Integrate a batch of members on the CPU.  Shapes as in include/lwfb90.h.
    Returns (daily [M, n_days+1, L] or None, status [M], counters [M,3])."""
    forcing = np.ascontiguousarray(forcing, dtype=np.float64)
    if forcing.ndim == 2:
        forcing = forcing[None]
    n_sites, n_rows, nf = forcing.shape
    scal = np.ascontiguousarray(scal, dtype=np.float64)
    lay = np.ascontiguousarray(lay, dtype=np.float64)
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    if scal.ndim == 1:
        scal, lay, u0 = scal[None], lay[None], u0[None]
    M = scal.shape[0]
    L = 9 * N + 51 if iso else 5 * N + 39
    assert u0.shape == (M, L), (u0.shape, (M, L))
    assert lay.shape[0] == M and lay.shape[2] == N
    som = np.ascontiguousarray(site_of_member, dtype=np.int32)
    assert som.shape == (M,)
    out = np.empty((M, n_days + 1, L)) if record else None
    status = np.zeros(M, dtype=np.int32)
    counters = np.zeros((M, 3), dtype=np.int64)
    lib().lwfb_oracle_run(N, int(iso), int(t0_day), int(n_days), n_sites, n_rows, _p(forcing), M, _p(som, _ip),
                          _p(scal), _p(lay), _p(u0), reltol, abstol, dt_init, dtmax, maxiters, int(n_threads),
                          _p(out) if record else None, _p(status, _ip), _p(counters, _lp))
    return out, status, counters


def run_member(N, iso, t0_day, n_days, forcing, scal, lay, u0, *, reltol=1e-5, abstol=1e-6, dt_init=1e-3,
               dtmax=60 / 60 / 24, maxiters=1_000_000, ntfli_init=None):
    """One member; optionally with warm du_NTFLI cache (re-run on the same problem).
    Returns (daily [n_days+1, L], status, counters[3], ntfli_final[N])."""
    forcing = np.ascontiguousarray(forcing, dtype=np.float64)
    n_rows = forcing.shape[0]
    L = 9 * N + 51 if iso else 5 * N + 39
    out = np.empty((n_days + 1, L))
    counters = np.zeros(3, dtype=np.int64)
    nt_out = np.empty(N)
    nt_in = None if ntfli_init is None else np.ascontiguousarray(ntfli_init, dtype=np.float64)
    st = lib().lwfb_oracle_run_member(N, int(iso), int(t0_day), int(n_days), n_rows, _p(forcing),
                                      _p(np.ascontiguousarray(scal, dtype=np.float64)),
                                      _p(np.ascontiguousarray(lay, dtype=np.float64)),
                                      _p(np.ascontiguousarray(u0, dtype=np.float64)),
                                      reltol, abstol, dt_init, dtmax, maxiters,
                                      None if nt_in is None else _p(nt_in), _p(out), _p(counters, _lp), _p(nt_out))
    return out, st, counters, nt_out


def eval_rhs(N, scal, lay, t, swati, gwat, frozen, ntfli_prev):
    sw = np.ascontiguousarray(swati, dtype=np.float64).copy()
    du = np.empty(N + 9)
    lib().lwfb_oracle_eval_rhs(N, _p(np.ascontiguousarray(scal, dtype=np.float64)),
                               _p(np.ascontiguousarray(lay, dtype=np.float64)), float(t), _p(sw), float(gwat),
                               _p(np.ascontiguousarray(frozen, dtype=np.float64)),
                               _p(np.ascontiguousarray(ntfli_prev, dtype=np.float64)), _p(du))
    return du, sw


def elem(fn: int, x, y=None):
    """Elementary function as compiled into the selected oracle build."""
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    y = np.zeros_like(x) if y is None else np.ascontiguousarray(y, dtype=np.float64).ravel()
    L = lib()
    if fn == 8:
        return np.array([L.lwfb_oracle_fastpow(a, b) for a, b in zip(x, y)])
    return np.array([L.lwfb_oracle_elem(fn, a, b) for a, b in zip(x, y)])


def setup_members(N, iso, base_scal, base_lay, base_u0, knobs, over_idx, profiles, psim_base, ilayer):
    """CPU twin of lwfb_setup_members (include/lwfb90_setup.h compiled for the host; always the "elem" build).
    Returns (scal [M, PS], lay [M, PL, N], u0 [M, L], status [M])."""
    build()
    L_ = _libs.get("setup") or C.CDLL(_SO_ELEM)
    _libs["setup"] = L_
    L_.lwfb_oracle_setup_members.restype = None
    L_.lwfb_oracle_setup_members.argtypes = [C.c_int64, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_int, _ip, C.c_int, _dp, _dp,
                                             C.c_int, _dp, _dp, _dp, _ip]
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    knobs = f(knobs)
    oi = np.ascontiguousarray(over_idx, dtype=np.int32)
    M = knobs.shape[0]
    PS, PL = len(base_scal), np.asarray(base_lay).shape[0]
    Ls = 9 * N + 51 if iso else 5 * N + 39
    profiles = f(profiles).reshape(-1, 2 * N + 1) if profiles is not None and len(profiles) else np.zeros((0, 2 * N + 1))
    scal, lay, u0 = np.empty((M, PS)), np.empty((M, PL, N)), np.empty((M, Ls))
    status = np.zeros(M, dtype=np.int32)
    L_.lwfb_oracle_setup_members(M, N, int(iso), _p(f(base_scal)), _p(f(base_lay)), _p(f(base_u0)), _p(knobs), len(oi),
                                 _p(oi, _ip) if len(oi) else None, profiles.shape[0], _p(profiles) if profiles.size else None,
                                 _p(f(psim_base)), int(ilayer), _p(scal), _p(lay), _p(u0), _p(status, _ip))
    return scal, lay, u0, status
