"""Host-side mirror of the reference's simulate! and result getters, on top of the C ABI.

Mirrors (does not copy) LWFBrook90.jl v0.9.9:
  simulate!               src/LWFBrook90.jl:628-655
  get_soil_idx            src/func_postprocess.jl:1025-1045
  get_soil_               src/func_postprocess.jl:1097-1159
  get_δsoil / get_amounts / get_δ / get_scalars   :1170-1298
  get_water_partitioning  :1310-1364
  get_daily_soilFluxes    test/03-regression-tests.jl:231-273 (helper of the regression tests)

All numerics of the simulation itself run in liblwfb90.so on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import datetime as _dt
from typing import Dict, Iterable, Optional, Sequence

import numpy as np

from . import capi
from . import layout as lo
from .ensemble import Ensemble
from .setup import DiscretizedSPAC, iso_prec_at

RETCODES = {0: "Success", 1: "Unstable", 2: "DtLessThanMin", 3: "MaxIters", 4: "RSSError", 5: "InvalidInput"}


class StateView:
    """One saved state vector with the reference's ComponentArray field names."""

    def __init__(self, rec: np.ndarray, ix: lo.StateIndex):
        self._r, self._ix = rec, ix

    def __getattr__(self, name):
        ix, r = self._ix, self._r
        if name in ("SWATI", "TRANI"):
            return _Triple(r, getattr(ix, name + "_mm"), getattr(ix, name + "_d18O"), getattr(ix, name + "_d2H"),
                           "mm" if name == "SWATI" else "mmday")
        if name in ("GWAT", "INTS", "INTR", "SNOW", "RWU", "XYLEM"):
            return _Triple(r, getattr(ix, name + "_mm"), getattr(ix, name + "_d18O"), getattr(ix, name + "_d2H"),
                           "mmday" if name == "RWU" else "mm")
        if name == "CC":
            return _Named({"MJm2": r[ix.CC]})
        if name == "SNOWLQ":
            return _Named({"mm": r[ix.SNOWLQ]})
        if name == "aux":
            return _Named({"θ": r[ix.aux_theta], "theta": r[ix.aux_theta], "ψ": r[ix.aux_psi], "K": r[ix.aux_K]})
        if name == "accum":
            return _Named({n: r[ix.accum0 + i] for i, n in enumerate(lo.ACCUM_NAMES)})
        raise AttributeError(name)


class _Named:
    def __init__(self, d):
        self.__dict__.update(d)
        self._d = d

    def __getitem__(self, k):
        return self._d[k]


class _Triple(_Named):
    def __init__(self, r, a, b, c, first):
        d = {first: r[a]}
        if b is not None:
            d["d18O"], d["d2H"] = r[b], r[c]
        super().__init__(d)


class Destats:
    def __init__(self, naccept, nreject, nf):
        self.naccept, self.nreject, self.nf = int(naccept), int(nreject), int(nf)


class ODESolution:
    """What the post-processing of the reference reads from an ODESolution
    (SURVEY 8b): .t, .u[i], sol(t), .retcode, .destats, .prob.p essentials."""

    def __init__(self, sim: DiscretizedSPAC, t: np.ndarray, records: np.ndarray, status: int, counters):
        self.t = np.asarray(t, float)
        self.records = records                       # [len(t), L]
        self.ix = lo.StateIndex(sim.N, sim.iso)
        self.retcode = RETCODES.get(int(status), f"status {status}")
        self.status = int(status)
        self.destats = Destats(*counters)
        self.stats = self.destats
        self.p_soil = sim.p_soil
        self.reference_date = sim.parametrizedSPAC.reference_date
        self.simulate_isotopes = sim.iso
        self._sim = sim

    @property
    def u(self):
        return [StateView(r, self.ix) for r in self.records]

    def __len__(self):
        return len(self.t)

    def __getitem__(self, i):
        return StateView(self.records[i], self.ix)

    def record_at(self, t: float) -> np.ndarray:
        """sol(t): the saved vector at a saved t, linear interpolation in between (the reference's
        solution is not dense because saveat is set)."""
        tt = self.t
        if t < tt[0] or t > tt[-1]:
            raise ValueError(f"t={t} outside the solution time span ({tt[0]}, {tt[-1]})")
        j = int(np.searchsorted(tt, t, side="left"))
        if tt[j] == t:
            return self.records[j]
        a, b = self.records[j - 1], self.records[j]
        w = (t - tt[j - 1]) / (tt[j] - tt[j - 1])
        return a + w * (b - a)

    def __call__(self, t: float) -> StateView:
        return StateView(self.record_at(float(t)), self.ix)

    # forcing lookups that the getters of the reference take from sol.prob.p
    def p_PREC(self, t):
        return float(self._sim.forcing[int(np.floor(t)), lo.F.PREC])

    def p_d18O_PREC(self, t):
        return iso_prec_at(self._sim.parametrizedSPAC.forcing["meteo_iso"], t, "delta18O_permil")

    def p_d2H_PREC(self, t):
        return iso_prec_at(self._sim.parametrizedSPAC.forcing["meteo_iso"], t, "delta2H_permil")


def _successful(status: int) -> bool:
    return status == 0


_SOLVE_KW = {"reltol": "reltol", "abstol": "abstol", "dt": "dt_init", "dtmax": "dtmax", "maxiters": "maxiters"}


def simulate_(s: DiscretizedSPAC, *, assert_retcode: bool = True, saveat=None, save_everystep=None, tspan=None,
              device: int = 0, **kwargs) -> None:
    """simulate!(s): integrates the discretised SPAC on the GPU and stores the solution in
    ``s.solution`` (``s.ODESolution``).

    The reference forwards its keyword arguments to solve() (src/LWFBrook90.jl:628-655).  Supported here:
    ``tspan`` (whole days), ``reltol``, ``abstol``, ``dt``, ``dtmax``, ``maxiters`` (forwarded to the integrator),
    ``saveat`` (whole days inside tspan: the library records the state at every day boundary).  NOT supported, and
    refused rather than answered with something else: ``save_everystep=True`` (the reference's default; the record
    of every accepted step is not kept on the GPU) is accepted only as the implicit default and means daily records
    here; an explicit ``save_everystep=True``, a ``saveat`` between day boundaries (the reference interpolates between
    its accepted steps, which would need them), and any other solve() keyword raise."""
    solver = {}
    for k in list(kwargs):
        if k in _SOLVE_KW:
            solver[_SOLVE_KW[k]] = kwargs.pop(k)
    if kwargs:
        raise TypeError(f"simulate_: unsupported solve() keyword(s) {sorted(kwargs)} (supported: tspan, saveat on day "
                        f"boundaries, {', '.join(sorted(_SOLVE_KW))})")
    if save_everystep:
        raise NotImplementedError("simulate_: save_everystep=True is not supported: the GPU integrator records the state "
                                  "at day boundaries only (saveat = t0:1:t1)")
    if saveat is not None:
        sa_chk = np.asarray(list(saveat), float)
        if np.any(sa_chk != np.floor(sa_chk)):
            raise NotImplementedError("simulate_: saveat must lie on day boundaries; values in between would have to be "
                                      "interpolated between the accepted steps of the solver, which are not recorded")
    t0, t1 = s.tspan if tspan is None else (float(tspan[0]), float(tspan[1]))
    if t0 != np.floor(t0) or t1 != np.floor(t1) or t1 <= t0:
        raise ValueError("tspan must have whole-day bounds")
    n_days = int(t1 - t0)
    h = capi.Handle(s.N, s.iso, int(t0), n_days, device=device, record_daily=True, **solver)
    try:
        h.set_sites(s.forcing)
        h.set_members([0], s.scal, s.lay, s.u0)
        h.run()
        rec = h.get_daily()[0]
        st, cnt = h.get_status()
    finally:
        h.close()
    t_all = t0 + np.arange(n_days + 1, dtype=float)
    if saveat is not None:
        sa = np.asarray(list(saveat), float)
        sol_full = ODESolution(s, t_all, rec, st[0], cnt[0])
        recs = np.stack([sol_full.record_at(x) for x in sa])
        s.solution = ODESolution(s, sa, recs, st[0], cnt[0])
    else:
        s.solution = ODESolution(s, t_all, rec, st[0], cnt[0])
    s.ODESolution_datetime = [s.parametrizedSPAC.reference_date + _dt.timedelta(seconds=float(np.floor(86400 * x)))
                              for x in s.solution.t]
    if assert_retcode and not _successful(int(st[0])):
        raise AssertionError(f"Problem with simulation: Return code of simulation was '{s.solution.retcode}'")


simulate = simulate_


class EnsembleResult:
    """Device-resident result of an ensemble run (one GPU)."""

    def __init__(self, ens: Ensemble, handle: capi.Handle):
        self.ens, self.handle = ens, handle
        self.status, self.counters = handle.get_status()

    def daily(self, first=0, count=None):
        return self.handle.get_daily(first, count)

    def final(self, first=0, count=None):
        return self.handle.get_final(first, count)

    def scores(self, obs_day, obs_layer, obs_theta=None, obs_d18o=None):
        return self.handle.scores(obs_day, obs_layer, obs_theta, obs_d18o)

    def water_partitioning(self, reference_date, period: str = "month"):
        """Ensemble form of get_water_partitioning (src/func_postprocess.jl:1310-1364), aggregated on the GPU: returns
        (labels, table) with labels = the (year, month) / (year,) of each period and table [M, n_periods, NWP] in
        layout.WP column order (sums; Swat: mean).  The daily record stays on the device."""
        dates = [reference_date + _dt.timedelta(days=float(t)) for t in range(self.ens.t0_day, self.ens.t0_day + self.handle.n_days + 1)]
        lab = [(d.year, d.month) if period == "month" else (d.year,) for d in dates]
        uniq = []
        for l in lab:
            if l not in uniq:
                uniq.append(l)
        group = np.array([uniq.index(l) for l in lab], dtype=np.int32)
        return uniq, self.handle.water_partitioning(group, len(uniq))

    def soil_state(self, days_to_read_out_d, *, depths_to_read_out_mm=None, layers=None):
        """Ensemble form of get_soil_ (src/func_postprocess.jl:1025-1045), computed on the GPU: [M, n_days, n_layers,
        NSV] in layout.SV column order for the given days (record indices relative to the start of the run) and either
        depths (mm, looked up in the shared grid as get_soil_idx does) or 0-based layer indices."""
        if layers is None:
            lower = np.cumsum(self.ens.lay[0, lo.L.THICK])
            layers = [int(np.nonzero(d <= lower)[0][0]) for d in depths_to_read_out_mm if 0 < d <= lower.max()]
        return self.handle.soil_state(np.asarray(days_to_read_out_d, dtype=np.int32), np.asarray(layers, dtype=np.int32))

    def close(self):
        self.handle.close()


def simulate_ensemble(ens: Ensemble, *, device: int = 0, record_daily: bool = True, n_days: Optional[int] = None,
                      handle: Optional[capi.Handle] = None) -> EnsembleResult:
    """Integrates all members of an ensemble on one GPU (the batched counterpart of simulate!)."""
    nd = ens.n_days if n_days is None else int(n_days)
    h = handle or capi.Handle(ens.N, ens.iso, ens.t0_day, nd, device=device, record_daily=record_daily)
    h.set_sites(ens.forcing)
    h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    h.run()
    return EnsembleResult(ens, h)


# ----------------------------------------------------------------------------- getters

def _sol(sim: DiscretizedSPAC) -> ODESolution:
    if sim.solution is None:
        raise AssertionError("Solution was not yet computed. Please simulate_(simulation)")
    return sim.solution


def get_soil_idx(sim: DiscretizedSPAC, depths_to_read_out_mm, only_valid_idxs=False):
    depths = list(depths_to_read_out_mm)
    assert all(d > 0 for d in depths)
    lower = np.cumsum(sim.p_soil.p_THICK)
    out = {}
    for d in depths:
        if d > lower.max():
            out[d] = 0
        else:
            out[d] = int(np.nonzero(d <= lower)[0][0]) + 1      # 1-based like the reference
    if only_valid_idxs:
        return [v for v in out.values() if v != 0]
    return out


_SYM = {"θ": "theta", "theta": "theta", "ψ": "psi", "psi": "psi", "δ18O": "d18O", "delta18O": "d18O",
        "δ2H": "d2H", "delta2H": "d2H", "W": "W", "SWATI": "SWATI", "K": "K"}


def get_soil_(symbols, sim: DiscretizedSPAC, *, depths_to_read_out_mm=None, days_to_read_out_d=None) -> Dict[str, np.ndarray]:
    """Table (dict of columns) of soil variables: rows = time steps, columns = '<symbol>_Lay<i>' or
    '<symbol>_<depth>mm'.  Symbols: θ/theta, ψ/psi, W, SWATI, K, δ18O/delta18O, δ2H/delta2H."""
    sol = _sol(sim)
    if isinstance(symbols, str):
        symbols = [symbols]
    tp = sol.t if days_to_read_out_d is None else np.atleast_1d(np.asarray(days_to_read_out_d, float))
    assert np.all(tp >= 0)
    recs = np.stack([sol.record_at(float(t)) for t in tp])
    ix = sol.ix
    SW = recs[:, ix.SWATI_mm]
    W, psi, _, theta, K = (np.stack(x) for x in zip(*(sim.p_soil.soilvar(r) for r in SW)))
    data = {"theta": theta, "psi": psi, "W": W, "SWATI": SW, "K": K}
    if sim.iso:
        data["d18O"], data["d2H"] = recs[:, ix.SWATI_d18O], recs[:, ix.SWATI_d2H]
    out = {"time": tp.copy()}
    idxs = None if depths_to_read_out_mm is None else get_soil_idx(sim, depths_to_read_out_mm)
    for sym in sorted(symbols):
        if sym not in _SYM:
            raise ValueError(f"Unknown symbol {sym} requested in `get_soil_()`")
        arr = data.get(_SYM[sym])
        if arr is None:
            arr = np.full_like(SW, np.nan)
        if idxs is None:
            for i in range(arr.shape[1]):
                out[f"{sym}_Lay{i + 1}"] = arr[:, i]
        else:
            for d, v in sorted(idxs.items()):
                if v != 0:
                    out[f"{sym}_{int(d)}mm"] = arr[:, v - 1]
    return out


def get_theta(sim, **kw):
    return get_soil_("θ", sim, **kw)


def get_psi(sim, **kw):
    return get_soil_("ψ", sim, **kw)


def get_deltasoil(sim: DiscretizedSPAC, *, depths_to_read_out_mm=None, days_to_read_out_d=None):
    """(d18O, d2H) matrices with soil layers as rows and time steps as columns (reference get_δsoil)."""
    sol = _sol(sim)
    assert sim.iso, "Provided solution did not simulate isotopes"
    tp = sol.t if days_to_read_out_d is None else np.atleast_1d(np.asarray(days_to_read_out_d, float))
    recs = np.stack([sol.record_at(float(t)) for t in tp])
    d18, d2 = recs[:, sol.ix.SWATI_d18O].T, recs[:, sol.ix.SWATI_d2H].T
    if depths_to_read_out_mm is not None:
        sel = [i - 1 for i in get_soil_idx(sim, depths_to_read_out_mm, only_valid_idxs=True)]
        d18, d2 = d18[sel], d2[sel]
    return {"d18O": d18, "d2H": d2}


def _get_scalars(comps, units, sim, days):
    sol = _sol(sim)
    tp = sol.t if days is None else np.atleast_1d(np.asarray(days, float))
    recs = np.stack([sol.record_at(float(t)) for t in tp])
    out = {"time": tp.copy()}
    ix = sol.ix
    for c, u in zip(comps, units):
        if c == "PREC":
            f = {"mm": sol.p_PREC, "mmday": sol.p_PREC, "d18O": sol.p_d18O_PREC, "d2H": sol.p_d2H_PREC}[u]
            out[f"PREC_{u}"] = np.array([f(float(t)) for t in tp])
    for c, u in zip(comps, units):
        if c == "PREC":
            continue
        if c in ("CC", "SNOWLQ"):
            col = getattr(ix, c)
        else:
            col = getattr(ix, f"{c}_{'mm' if u in ('mm', 'mmday') else u}")
        out[f"{c}_{u}"] = recs[:, col]
    return out


def get_amounts(sim: DiscretizedSPAC, *, days_to_read_out_d=None):
    return _get_scalars(["GWAT", "INTS", "INTR", "SNOW", "CC", "SNOWLQ"], ["mm", "mm", "mm", "mm", "MJm2", "mm"],
                        sim, days_to_read_out_d)


def get_delta(sim: DiscretizedSPAC, *, days_to_read_out_d=None):
    comps = ["PREC", "GWAT", "INTS", "INTR", "SNOW", "RWU", "XYLEM"]
    a = _get_scalars(comps, ["d18O"] * 7, sim, days_to_read_out_d)
    b = _get_scalars(comps, ["d2H"] * 7, sim, days_to_read_out_d)
    b.pop("time")
    return {**a, **b}


get_δ = get_delta


def get_daily_soilFluxes(sim: DiscretizedSPAC):
    """The 30 daily flux / storage series the reference's regression tests compare."""
    sol = _sol(sim)
    t_out = np.arange(sol.t.min(), sol.t.max() + 0.5, 1.0)
    recs = np.stack([sol.record_at(float(t)) for t in t_out])
    # column order of the reference's function (test/03-regression-tests.jl:231-273), not the state order
    names = ["cum_d_prec", "cum_d_rfal", "cum_d_sfal", "cum_d_rint", "cum_d_sint", "cum_d_rthr", "cum_d_sthr", "cum_d_rsno",
             "cum_d_rnet", "cum_d_smlt", "cum_d_evap", "cum_d_tran", "cum_d_irvp", "cum_d_isvp", "cum_d_slvp", "cum_d_snvp",
             "cum_d_pint", "cum_d_ptran", "flow", "seep", "srfl", "slfl", "byfl", "dsfl", "gwfl", "vrfln", "StorageSWAT",
             "StorageWATER", "BALERD_SWAT", "BALERD_total"]
    assert sorted(names) == sorted(n for n in lo.ACCUM_NAMES if n != "cum_d_pslvp")
    out = {"time": [sol.reference_date + _dt.timedelta(days=float(t)) for t in t_out]}
    for n in names:
        out[n] = recs[:, sol.ix.acc(n)]
    return out


def get_water_partitioning(sim: DiscretizedSPAC):
    """(daily, monthly, yearly) tables of the water partitioning; needs a daily-saved solution."""
    sol = _sol(sim)
    if not np.all(np.diff(sol.t) == 1.0):
        raise AssertionError("Solution is not computed with daily output resolution. Provide saveat in daily resolution.")
    ix = sol.ix
    A = lambda n: sol.records[:, ix.acc(n)]
    dates = [sol.reference_date + _dt.timedelta(days=float(t)) for t in sol.t]
    daily = {
        "date": dates, "year": np.array([d.year for d in dates]), "month": np.array([d.month for d in dates]),
        "ETa": A("cum_d_evap"), "Esoil": A("cum_d_slvp"), "Esnow": A("cum_d_snvp"),
        "Einterception": A("cum_d_irvp") + A("cum_d_isvp"), "Ta": A("cum_d_tran"), "Precip": A("cum_d_prec"),
        "Td": A("cum_d_ptran") - A("cum_d_tran"), "D": -A("vrfln"), "R": -(A("srfl") + A("byfl") + A("dsfl")),
        "Swat": A("StorageSWAT")}
    sums = ["ETa", "Esoil", "Esnow", "Einterception", "Ta", "Precip", "Td", "D", "R"]

    def group(keys):
        lab = list(zip(*(daily[k] for k in keys)))
        uniq = []
        for l in lab:
            if l not in uniq:
                uniq.append(l)
        out = {k: np.array([u[i] for u in uniq]) for i, k in enumerate(keys)}
        lab_arr = np.array([uniq.index(l) for l in lab])
        out["nrow"] = np.array([(lab_arr == g).sum() for g in range(len(uniq))])
        for c in sums:
            out[c] = np.array([daily[c][lab_arr == g].sum() for g in range(len(uniq))])
        out["Swat"] = np.array([daily["Swat"][lab_arr == g].mean() for g in range(len(uniq))])
        return out
    return daily, group(["year", "month"]), group(["year"])
