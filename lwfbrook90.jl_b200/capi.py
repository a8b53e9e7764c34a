"""ctypes binding of the C ABI (include/lwfb90.h) of liblwfb90.so.

This is the Python equivalent of the Julia `ccall` shim shown in INTEGRATION.md.  There is no CPU
fallback: if the CUDA library is missing or no GPU is present the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import layout as lo

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LWFB_LIB", os.path.join(_HERE, "liblwfb90.so"))

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


class LwfbOpts(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("simulate_isotopes", C.c_int32), ("t0_day", C.c_int32),
                ("n_days", C.c_int32), ("reltol", C.c_double), ("abstol", C.c_double),
                ("dt_init", C.c_double), ("dtmax", C.c_double), ("maxiters", C.c_int64),
                ("device", C.c_int32), ("record_daily", C.c_int32)]


class LwfbError(RuntimeError):
    pass


_lib = None

# every symbol include/lwfb90.h declares
SYMBOLS = ["lwfb_default_opts", "lwfb_create", "lwfb_destroy", "lwfb_last_error", "lwfb_set_sites",
           "lwfb_set_members", "lwfb_run", "lwfb_run_async", "lwfb_sync", "lwfb_get_daily", "lwfb_get_final",
           "lwfb_get_status", "lwfb_scores", "lwfb_scores_device", "lwfb_eval_rhs", "lwfb_last_run_stats", "lwfb_elem_eval",
           "lwfb_fp64_peak", "lwfb_debug_get_block", "lwfb_debug_force_legacy", "lwfb_set_placement",
           "lwfb_water_partitioning", "lwfb_soil_state", "lwfb_set_observations", "lwfb_setup_members",
           "lwfb_set_ntfli0", "lwfb_get_ntfli",
           "lwfb_multi_create", "lwfb_multi_destroy", "lwfb_multi_last_error", "lwfb_multi_n_shards", "lwfb_multi_shard",
           "lwfb_multi_set_sites", "lwfb_multi_set_members", "lwfb_multi_setup_members", "lwfb_multi_set_observations",
           "lwfb_multi_run", "lwfb_multi_allgather_scores", "lwfb_multi_get_status", "lwfb_multi_get_final",
           "lwfb_multi_last_run_ms"]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LwfbError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.lwfb_default_opts.argtypes = [C.POINTER(LwfbOpts)]; L.lwfb_default_opts.restype = None
        L.lwfb_create.argtypes = [C.POINTER(vp), C.POINTER(LwfbOpts)]; L.lwfb_create.restype = C.c_int
        L.lwfb_destroy.argtypes = [vp]; L.lwfb_destroy.restype = None
        L.lwfb_last_error.argtypes = [vp]; L.lwfb_last_error.restype = C.c_char_p
        L.lwfb_set_sites.argtypes = [vp, C.c_int32, C.c_int32, _dp]; L.lwfb_set_sites.restype = C.c_int
        L.lwfb_set_members.argtypes = [vp, C.c_int64, _ip, _dp, _dp, _dp]; L.lwfb_set_members.restype = C.c_int
        L.lwfb_run.argtypes = [vp, vp]; L.lwfb_run.restype = C.c_int
        L.lwfb_run_async.argtypes = [vp, vp]; L.lwfb_run_async.restype = C.c_int
        L.lwfb_sync.argtypes = [vp]; L.lwfb_sync.restype = C.c_int
        L.lwfb_get_daily.argtypes = [vp, C.c_int64, C.c_int64, _dp]; L.lwfb_get_daily.restype = C.c_int
        L.lwfb_get_final.argtypes = [vp, C.c_int64, C.c_int64, _dp]; L.lwfb_get_final.restype = C.c_int
        L.lwfb_get_status.argtypes = [vp, _ip, _lp]; L.lwfb_get_status.restype = C.c_int
        L.lwfb_scores.argtypes = [vp, C.c_int32, _ip, C.c_int32, _ip, _dp, _dp, _dp]; L.lwfb_scores.restype = C.c_int
        L.lwfb_scores_device.argtypes = [vp, C.c_int32, _ip, C.c_int32, _ip, _dp, _dp, C.POINTER(vp)]
        L.lwfb_scores_device.restype = C.c_int
        L.lwfb_set_observations.argtypes = [vp, C.c_int32, _ip, C.c_int32, _ip, _dp, _dp]; L.lwfb_set_observations.restype = C.c_int
        L.lwfb_setup_members.argtypes = [vp, C.c_int64, _ip, _dp, _dp, _dp, _dp, C.c_int32, _ip, C.c_int32, _dp, _dp, C.c_int32, _ip]
        L.lwfb_setup_members.restype = C.c_int
        L.lwfb_set_ntfli0.argtypes = [vp, _dp]; L.lwfb_set_ntfli0.restype = C.c_int
        L.lwfb_get_ntfli.argtypes = [vp, _dp]; L.lwfb_get_ntfli.restype = C.c_int
        L.lwfb_multi_create.argtypes = [C.POINTER(vp), C.POINTER(LwfbOpts), C.c_int32, _ip]; L.lwfb_multi_create.restype = C.c_int
        L.lwfb_multi_destroy.argtypes = [vp]; L.lwfb_multi_destroy.restype = None
        L.lwfb_multi_last_error.argtypes = [vp]; L.lwfb_multi_last_error.restype = C.c_char_p
        L.lwfb_multi_n_shards.argtypes = [vp]; L.lwfb_multi_n_shards.restype = C.c_int32
        L.lwfb_multi_shard.argtypes = [vp, C.c_int32]; L.lwfb_multi_shard.restype = vp
        L.lwfb_multi_set_sites.argtypes = [vp, C.c_int32, C.c_int32, _dp]; L.lwfb_multi_set_sites.restype = C.c_int
        L.lwfb_multi_set_members.argtypes = [vp, C.c_int64, _ip, _dp, _dp, _dp, _dp]; L.lwfb_multi_set_members.restype = C.c_int
        L.lwfb_multi_setup_members.argtypes = [vp, C.c_int64, _dp, _dp, _dp, _dp, C.c_int32, _ip, C.c_int32, _dp, _dp, C.c_int32, _dp]
        L.lwfb_multi_setup_members.restype = C.c_int
        L.lwfb_multi_set_observations.argtypes = [vp, C.c_int32, _ip, C.c_int32, _ip, _dp, _dp]; L.lwfb_multi_set_observations.restype = C.c_int
        L.lwfb_multi_run.argtypes = [vp]; L.lwfb_multi_run.restype = C.c_int
        L.lwfb_multi_allgather_scores.argtypes = [vp, C.c_int32, _ip, C.c_int32, _ip, _dp, _dp, C.POINTER(vp), _dp]
        L.lwfb_multi_allgather_scores.restype = C.c_int
        L.lwfb_multi_get_status.argtypes = [vp, _ip, _lp, _ip]; L.lwfb_multi_get_status.restype = C.c_int
        L.lwfb_multi_get_final.argtypes = [vp, _dp]; L.lwfb_multi_get_final.restype = C.c_int
        L.lwfb_multi_last_run_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]; L.lwfb_multi_last_run_ms.restype = C.c_int
        L.lwfb_eval_rhs.argtypes = [vp, C.c_double, _dp, _dp, _dp, _dp, _dp]; L.lwfb_eval_rhs.restype = C.c_int
        L.lwfb_last_run_stats.argtypes = [vp, _ip, C.POINTER(C.c_float)]; L.lwfb_last_run_stats.restype = C.c_int
        L.lwfb_elem_eval.argtypes = [C.c_int32, C.c_int32, C.c_int64, _dp, _dp, _dp]; L.lwfb_elem_eval.restype = C.c_int
        L.lwfb_water_partitioning.argtypes = [vp, C.c_int32, _ip, _dp]; L.lwfb_water_partitioning.restype = C.c_int
        L.lwfb_soil_state.argtypes = [vp, C.c_int32, _ip, C.c_int32, _ip, _dp]; L.lwfb_soil_state.restype = C.c_int
        _lib = L
    return _lib


def fp64_peak(device: int = 0, reps: int = 5) -> float:
    """Measured DFMA peak of the device in TFLOP/s."""
    out = C.c_double()
    L = lib()
    L.lwfb_fp64_peak.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    L.lwfb_fp64_peak.restype = C.c_int
    rc = L.lwfb_fp64_peak(device, reps, C.byref(out))
    if rc != 0:
        raise LwfbError(f"lwfb_fp64_peak failed with code {rc}")
    return float(out.value)


def elem_eval(fn: int, x, y=None, device: int = 0):
    """Evaluate an elementary function of include/lwfb90_elem.h on the GPU."""
    x = _f64(x).ravel()
    y = np.zeros_like(x) if y is None else _f64(y).ravel()
    out = np.empty_like(x)
    rc = lib().lwfb_elem_eval(device, fn, x.size, _p(x), _p(y), _p(out))
    if rc != 0:
        raise LwfbError(f"lwfb_elem_eval failed with code {rc}")
    return out


def _p(a, t=_dp):
    return a.ctypes.data_as(t)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Handle:
    """Owns one lwfb_handle (device memory of one ensemble on one GPU)."""

    def __init__(self, N: int, iso: bool, t0_day: int, n_days: int, *, device: int = 0, record_daily: bool = True,
                 reltol=1e-5, abstol=1e-6, dt_init=1e-3, dtmax=60 / 60 / 24, maxiters=1_000_000):
        self._L = lib()
        o = LwfbOpts()
        self._L.lwfb_default_opts(C.byref(o))
        o.n_layers, o.simulate_isotopes, o.t0_day, o.n_days = int(N), int(bool(iso)), int(t0_day), int(n_days)
        o.reltol, o.abstol, o.dt_init, o.dtmax, o.maxiters = reltol, abstol, dt_init, dtmax, int(maxiters)
        o.device, o.record_daily = int(device), int(bool(record_daily))
        self.N, self.iso, self.t0_day, self.n_days = int(N), bool(iso), int(t0_day), int(n_days)
        self.L = lo.state_len(self.N, self.iso)
        self.M = 0
        self._h = C.c_void_p()
        rc = self._L.lwfb_create(C.byref(self._h), C.byref(o))
        if rc != 0:
            self._h = C.c_void_p()
            raise LwfbError(f"lwfb_create failed with code {rc} (is a CUDA device available?)")

    def _ck(self, rc):
        if rc != 0:
            msg = self._L.lwfb_last_error(self._h)
            raise LwfbError(f"liblwfb90 error {rc}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.lwfb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_sites(self, forcing):
        f = _f64(forcing)
        if f.ndim == 2:
            f = f[None]
        assert f.shape[2] == lo.NF
        self._ck(self._L.lwfb_set_sites(self._h, f.shape[0], f.shape[1], _p(f)))

    def set_members(self, site_of_member, scal, lay, u0):
        scal, lay, u0 = _f64(scal), _f64(lay), _f64(u0)
        if scal.ndim == 1:
            scal, lay, u0 = scal[None], lay[None], u0[None]
        M = scal.shape[0]
        assert scal.shape == (M, lo.NPS) and lay.shape == (M, lo.NPL, self.N) and u0.shape == (M, self.L)
        som = np.ascontiguousarray(site_of_member, dtype=np.int32).reshape(M)
        self._ck(self._L.lwfb_set_members(self._h, M, _p(som, _ip), _p(scal), _p(lay), _p(u0)))
        self.M = M

    def setup_members(self, kb, site_of_member=None):
        """Batched setup ON THE DEVICE (lwfb_setup_members): kb is an ensemble.KnobBlock - the base member and one row of
        knob values per member; the member blocks are built in HBM.  Returns the number of invalid members."""
        knobs = _f64(kb.knobs)
        M = knobs.shape[0]
        bs, bl, bu = _f64(kb.base_scal), _f64(kb.base_lay), _f64(kb.base_u0)
        assert bs.shape == (lo.NPS,) and bl.shape == (lo.NPL, self.N) and bu.shape == (self.L,)
        oi = np.ascontiguousarray(kb.over_idx, dtype=np.int32)
        assert knobs.shape[1] == 8 + len(oi)
        pr = _f64(kb.profiles).reshape(-1, 2 * self.N + 1)
        ps = _f64(kb.psim_base)
        som = None if site_of_member is None else np.ascontiguousarray(site_of_member, dtype=np.int32).reshape(M)
        n_inv = C.c_int32()
        self._ck(self._L.lwfb_setup_members(self._h, M, None if som is None else _p(som, _ip), _p(bs), _p(bl), _p(bu), _p(knobs),
                                            len(oi), _p(oi, _ip) if len(oi) else None, pr.shape[0],
                                            _p(pr) if pr.size else None, _p(ps), int(kb.ilayer), C.byref(n_inv)))
        self.M = M
        return int(n_inv.value)

    def get_block(self, which: int):
        """Testing aid: copy of a device-resident member block (0 scal, 1 lay, 2 u0)."""
        shape = {0: (self.M, lo.NPS), 1: (self.M, lo.NPL, self.N), 2: (self.M, self.L)}[which]
        out = np.empty(shape)
        self._L.lwfb_debug_get_block.argtypes = [C.c_void_p, C.c_int32, _dp]
        self._ck(self._L.lwfb_debug_get_block(self._h, int(which), _p(out)))
        return out

    def set_ntfli0(self, ntfli0=None):
        """du_NTFLI cache that the first ITER of the following runs sees ([M, N]); None = a fresh problem (NaN)."""
        if ntfli0 is None:
            self._ck(self._L.lwfb_set_ntfli0(self._h, None))
            return
        a = _f64(ntfli0).reshape(self.M, self.N)
        self._ck(self._L.lwfb_set_ntfli0(self._h, _p(a)))

    def get_ntfli(self):
        """du_NTFLI as the last run left it ([M, N])."""
        out = np.empty((self.M, self.N))
        self._ck(self._L.lwfb_get_ntfli(self._h, _p(out)))
        return out

    def run(self, stream=None):
        self._ck(self._L.lwfb_run(self._h, C.c_void_p(stream) if stream else None))

    def run_async(self, stream=None):
        self._ck(self._L.lwfb_run_async(self._h, C.c_void_p(stream) if stream else None))

    def sync(self):
        self._ck(self._L.lwfb_sync(self._h))

    def get_daily(self, first=0, count=None, out=None):
        count = self.M - first if count is None else count
        shape = (count, self.n_days + 1, self.L)
        if out is None:
            out = np.empty(shape)
        elif not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.shape == shape
                  and out.flags.c_contiguous and out.flags.writeable):
            raise LwfbError(f"get_daily(out=...): need a writeable C-contiguous float64 array of shape {shape}")
        self._ck(self._L.lwfb_get_daily(self._h, first, count, _p(out)))
        return out

    def get_final(self, first=0, count=None):
        count = self.M - first if count is None else count
        out = np.empty((count, self.L))
        self._ck(self._L.lwfb_get_final(self._h, first, count, _p(out)))
        return out

    def get_status(self):
        st = np.empty(self.M, dtype=np.int32)
        cnt = np.empty((self.M, lo.NCOUNTERS), dtype=np.int64)
        self._ck(self._L.lwfb_get_status(self._h, _p(st, _ip), _p(cnt, _lp)))
        return st, cnt

    def set_observations(self, obs_day=None, obs_layer=None, obs_theta=None, obs_d18o=None):
        """Register the observations that the following runs score on the fly (needed with record_daily=False);
        afterwards scores() / scores_device() without arguments return the result.  No arguments: clear."""
        if obs_day is None:
            self._ck(self._L.lwfb_set_observations(self._h, 0, None, 0, None, None, None))
            return
        od = np.ascontiguousarray(obs_day, dtype=np.int32)
        ol = np.ascontiguousarray(obs_layer, dtype=np.int32)
        th = None if obs_theta is None else _f64(obs_theta).reshape(len(od), len(ol))
        d18 = None if obs_d18o is None else _f64(obs_d18o).reshape(len(od), len(ol))
        self._ck(self._L.lwfb_set_observations(self._h, len(od), _p(od, _ip), len(ol), _p(ol, _ip),
                                               None if th is None else _p(th), None if d18 is None else _p(d18)))

    def scores(self, obs_day=None, obs_layer=None, obs_theta=None, obs_d18o=None):
        if obs_day is None:           # the sums accumulated during the run (set_observations)
            out = np.empty((self.M, lo.NSCORES))
            self._ck(self._L.lwfb_scores(self._h, 0, None, 0, None, None, None, _p(out)))
            return out
        od = np.ascontiguousarray(obs_day, dtype=np.int32)
        ol = np.ascontiguousarray(obs_layer, dtype=np.int32)
        th = None if obs_theta is None else _f64(obs_theta).reshape(len(od), len(ol))
        d18 = None if obs_d18o is None else _f64(obs_d18o).reshape(len(od), len(ol))
        out = np.empty((self.M, lo.NSCORES))
        self._ck(self._L.lwfb_scores(self._h, len(od), _p(od, _ip), len(ol), _p(ol, _ip),
                                     None if th is None else _p(th), None if d18 is None else _p(d18), _p(out)))
        return out

    def scores_device(self, obs_day=None, obs_layer=None, obs_theta=None, obs_d18o=None) -> int:
        """Like scores() but leaves the [M][NSCORES] block on the GPU and returns its device address."""
        if obs_day is None:
            ptr = C.c_void_p()
            self._ck(self._L.lwfb_scores_device(self._h, 0, None, 0, None, None, None, C.byref(ptr)))
            return int(ptr.value)
        od = np.ascontiguousarray(obs_day, dtype=np.int32)
        ol = np.ascontiguousarray(obs_layer, dtype=np.int32)
        th = None if obs_theta is None else _f64(obs_theta).reshape(len(od), len(ol))
        d18 = None if obs_d18o is None else _f64(obs_d18o).reshape(len(od), len(ol))
        ptr = C.c_void_p()
        self._ck(self._L.lwfb_scores_device(self._h, len(od), _p(od, _ip), len(ol), _p(ol, _ip),
                                            None if th is None else _p(th), None if d18 is None else _p(d18),
                                            C.byref(ptr)))
        return int(ptr.value)

    def water_partitioning(self, group_of_record, n_groups=None):
        """Per member and period: sums of ETa, Esoil, Esnow, Einterception, Ta, Precip, Td, D, R and the mean of Swat
        (lo.WP columns), aggregated on the GPU from the daily record.  group_of_record [n_days+1]: period index of every
        daily record (< 0: skip).  Returns [M, n_groups, NWP]."""
        g = np.ascontiguousarray(group_of_record, dtype=np.int32)
        if g.shape != (self.n_days + 1,):
            raise LwfbError(f"group_of_record must have n_days + 1 = {self.n_days + 1} entries")
        ng = int(g.max()) + 1 if n_groups is None else int(n_groups)
        out = np.empty((self.M, ng, lo.NWP))
        self._ck(self._L.lwfb_water_partitioning(self._h, ng, _p(g, _ip), _p(out)))
        return out

    def soil_state(self, rec_idx, layer_idx):
        """theta, psi, W, SWATI, K, delta18O, delta2H (lo.SV columns) of the given daily records and 0-based layers for
        every member, computed on the GPU.  Returns [M, len(rec_idx), len(layer_idx), NSV]."""
        r = np.ascontiguousarray(rec_idx, dtype=np.int32).ravel()
        l = np.ascontiguousarray(layer_idx, dtype=np.int32).ravel()
        out = np.empty((self.M, len(r), len(l), lo.NSV))
        self._ck(self._L.lwfb_soil_state(self._h, len(r), _p(r, _ip), len(l), _p(l, _ip), _p(out)))
        return out

    def eval_rhs(self, t, swati, gwat, frozen, ntfli_prev):
        sw, gw, fr, nt = _f64(swati), _f64(gwat), _f64(frozen), _f64(ntfli_prev)
        du = np.empty((self.M, self.N + 9))
        self._ck(self._L.lwfb_eval_rhs(self._h, float(t), _p(sw), _p(gw), _p(fr), _p(nt), _p(du)))
        return du

    def force_legacy(self, on=True):
        """Testing aid: True / 1 = the thread-per-member kernel instead of the lane-per-layer kernel; 2 = the lane kernel's
        general instantiation where a lean one would be picked; False / 0 = defaults."""
        self._L.lwfb_debug_force_legacy.argtypes = [C.c_void_p, C.c_int32]
        self._ck(self._L.lwfb_debug_force_legacy(self._h, int(on)))

    def set_placement(self, mode: int):
        """0: one launch, members in the given order; 1 (default): work quanta with regrouping by progress (see
        lwfb_set_placement in include/lwfb90.h); 2: the same also for ensembles that do not fill the device."""
        self._L.lwfb_set_placement.argtypes = [C.c_void_p, C.c_int32]
        self._ck(self._L.lwfb_set_placement(self._h, int(mode)))

    def last_run_stats(self):
        n = C.c_int32()
        ms = C.c_float()
        self._ck(self._L.lwfb_last_run_stats(self._h, C.byref(n), C.byref(ms)))
        return int(n.value), float(ms.value)


class MultiHandle:
    """Several GPUs of one host behind one handle (lwfb_multi_* of include/lwfb90.h): shards the members over the listed
    devices (a device may be listed more than once), runs the shards concurrently, all-gathers the scores peer to peer.
    No torch, no NCCL."""

    def __init__(self, N: int, iso: bool, t0_day: int, n_days: int, devices, *, record_daily: bool = False,
                 reltol=1e-5, abstol=1e-6, dt_init=1e-3, dtmax=60 / 60 / 24, maxiters=1_000_000):
        self._L = lib()
        o = LwfbOpts()
        self._L.lwfb_default_opts(C.byref(o))
        o.n_layers, o.simulate_isotopes, o.t0_day, o.n_days = int(N), int(bool(iso)), int(t0_day), int(n_days)
        o.reltol, o.abstol, o.dt_init, o.dtmax, o.maxiters = reltol, abstol, dt_init, dtmax, int(maxiters)
        o.record_daily = int(bool(record_daily))
        self.N, self.iso, self.n_days = int(N), bool(iso), int(n_days)
        self.L = lo.state_len(self.N, self.iso)
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        self.n_shards = len(dev)
        self.M = 0
        self._h = C.c_void_p()
        rc = self._L.lwfb_multi_create(C.byref(self._h), C.byref(o), len(dev), _p(dev, _ip))
        if rc != 0:
            self._h = C.c_void_p()
            raise LwfbError(f"lwfb_multi_create failed with code {rc}")

    def _ck(self, rc):
        if rc != 0:
            msg = self._L.lwfb_multi_last_error(self._h)
            raise LwfbError(f"liblwfb90 (multi) error {rc}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.lwfb_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_sites(self, forcing):
        f = _f64(forcing)
        if f.ndim == 2:
            f = f[None]
        self._ck(self._L.lwfb_multi_set_sites(self._h, f.shape[0], f.shape[1], _p(f)))

    def set_members(self, site_of_member, scal, lay, u0, cost=None):
        scal, lay, u0 = _f64(scal), _f64(lay), _f64(u0)
        M = scal.shape[0]
        som = np.ascontiguousarray(site_of_member, dtype=np.int32).reshape(M)
        c = None if cost is None else _f64(cost).reshape(M)
        self._ck(self._L.lwfb_multi_set_members(self._h, M, _p(som, _ip), _p(scal), _p(lay), _p(u0), None if c is None else _p(c)))
        self.M = M

    def setup_members(self, kb, cost=None):
        knobs = _f64(kb.knobs)
        M = knobs.shape[0]
        oi = np.ascontiguousarray(kb.over_idx, dtype=np.int32)
        pr = _f64(kb.profiles).reshape(-1, 2 * self.N + 1)
        c = None if cost is None else _f64(cost).reshape(M)
        self._ck(self._L.lwfb_multi_setup_members(self._h, M, _p(_f64(kb.base_scal)), _p(_f64(kb.base_lay)), _p(_f64(kb.base_u0)),
                                                  _p(knobs), len(oi), _p(oi, _ip) if len(oi) else None, pr.shape[0],
                                                  _p(pr) if pr.size else None, _p(_f64(kb.psim_base)), int(kb.ilayer),
                                                  None if c is None else _p(c)))
        self.M = M

    def set_observations(self, obs_day, obs_layer, obs_theta=None, obs_d18o=None):
        od = np.ascontiguousarray(obs_day, dtype=np.int32)
        ol = np.ascontiguousarray(obs_layer, dtype=np.int32)
        th = None if obs_theta is None else _f64(obs_theta).reshape(len(od), len(ol))
        d18 = None if obs_d18o is None else _f64(obs_d18o).reshape(len(od), len(ol))
        self._ck(self._L.lwfb_multi_set_observations(self._h, len(od), _p(od, _ip), len(ol), _p(ol, _ip),
                                                     None if th is None else _p(th), None if d18 is None else _p(d18)))

    def run(self):
        self._ck(self._L.lwfb_multi_run(self._h))

    def allgather_scores(self):
        """Scores of the whole ensemble (from the sums of the run), gathered on every device; returns the host copy and
        the per-shard device pointers."""
        out = np.empty((self.M, lo.NSCORES))
        ptrs = (C.c_void_p * self.n_shards)()
        self._ck(self._L.lwfb_multi_allgather_scores(self._h, 0, None, 0, None, None, None, ptrs, _p(out)))
        return out, [int(p) for p in ptrs]

    def get_status(self):
        st = np.empty(self.M, dtype=np.int32)
        cnt = np.empty((self.M, lo.NCOUNTERS), dtype=np.int64)
        sh = np.empty(self.M, dtype=np.int32)
        self._ck(self._L.lwfb_multi_get_status(self._h, _p(st, _ip), _p(cnt, _lp), _p(sh, _ip)))
        return st, cnt, sh

    def get_final(self):
        out = np.empty((self.M, self.L))
        self._ck(self._L.lwfb_multi_get_final(self._h, _p(out)))
        return out

    def last_run_ms(self):
        ms = (C.c_float * self.n_shards)()
        mx = C.c_float()
        self._ck(self._L.lwfb_multi_last_run_ms(self._h, ms, C.byref(mx)))
        return [float(x) for x in ms], float(mx.value)
