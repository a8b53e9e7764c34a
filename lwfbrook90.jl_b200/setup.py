"""Host-side twin of the reference's ``setup`` / ``remakeSPAC``: turns a parametrised SPAC into the
flat blocks the C ABI consumes (include/lwfb90.h).

Follows (does not copy) LWFBrook90.jl v0.9.9:
  setup                                src/LWFBrook90.jl:365-587
  remakeSPAC / remake_*                src/LWFBrook90.jl:214-345
  define_LWFB90_p                      src/func_DiffEq_definition_p.jl:12-586
  interpolate_meteo                    src/func_DiffEq_definition_p.jl:637-752
  generate_canopy_timeseries_relative  src/func_DiffEq_definition_p.jl:759-803
  make_absolute_from_relative          src/func_DiffEq_definition_p.jl:816-833
  define_LWFB90_u0 / init_LWFB90_u0!   src/func_DiffEq_definition_u0.jl:6-139
  WAT.INFPAR                           src/module_WAT.jl:59-79
  SUN.EQUIVSLP                         src/module_SUN.jl:73-85
"""
from __future__ import annotations

import copy
import datetime as _dt
import math
import warnings
from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import numpy as np

from . import layout as lo
from .discretize import (map_soil_horizons_to_discretization, overwrite_IC, overwrite_rootden,
                         refine_soil_discretization, root_growth_tini)
from .inputs import SPAC, doy_of, month_of
from .kpt import FTheta_MvG, FWETNES, SoilPar

P_PI = 3.1416  # src/module_CONSTANTS.jl:17 (sic)


@dataclass
class DiscretizedSPAC:
    """A discretised SPAC ready for simulation (reference: struct DiscretizedSPAC,
    src/LWFBrook90.jl:89-97).  Instead of an ODEProblem it carries the flat blocks of the C ABI."""
    parametrizedSPAC: SPAC
    N: int
    iso: bool
    tspan: tuple
    scal: np.ndarray          # [LWFB_PS]
    lay: np.ndarray           # [LWFB_PL, N]
    u0: np.ndarray            # [L]
    forcing: np.ndarray       # [n_rows, LWFB_F]
    p_soil: SoilPar
    ILAYER: int = 0           # number of layers receiving infiltration (IDEPTH)
    QLAYER: int = 0           # number of layers counted for source-area flow (QDEPTH)
    solution: Any = None      # filled by simulate_ (ODESolution-like object)

    @property
    def t0_day(self) -> int:
        return int(self.tspan[0])

    @property
    def n_days(self) -> int:
        return int(self.tspan[1] - self.tspan[0])

    @property
    def ODESolution(self):
        return self.solution


# ----------------------------------------------------------------------------- forcing

def iso_prec_at(meteo_iso, t: float, which: str) -> float:
    """δ of precipitation at time t: Gridded(Constant{Next}) on sample days + 0.5 with flat
    extrapolation (reference src/func_DiffEq_definition_p.jl:735-736)."""
    knots = meteo_iso["days"] + 12 / 24
    i = int(np.searchsorted(knots, t, side="left"))
    i = min(i, len(knots) - 1)
    return float(meteo_iso[which][i])


def canopy_relative(canopy_evolution, days, reference_date):
    """Relative canopy time series in percent (reference :759-803). ``days`` has n_days+1 entries."""
    if canopy_evolution.get("_is_table", False):
        return {k: np.asarray(canopy_evolution[k], float)
                for k in ("days", "DENSEF_rel", "HEIGHT_rel", "LAI_rel", "SAI_rel")}
    n = len(days)
    lai = canopy_evolution["LAI_rel"]
    knots = [lai["DOY_Bstart"], lai["DOY_Bstart"] + lai["Bduration"],
             lai["DOY_Cstart"], lai["DOY_Cstart"] + lai["Cduration"]]
    assert 1 <= knots[0] <= knots[1] <= knots[2] <= knots[3] <= 366, "DOY of LAI are not ordered."
    values = [lai["LAI_perc_CtoB"], lai["LAI_perc_BtoC"], lai["LAI_perc_BtoC"], lai["LAI_perc_CtoB"]]
    doys = np.array([(reference_date + _dt.timedelta(days=int(d))).timetuple().tm_yday for d in days], float)
    lai_rel = _linear_interp_flat(np.array(knots, float), np.array(values, float), doys)
    return {"days": np.asarray(days, float),
            "DENSEF_rel": np.full(n, float(canopy_evolution["DENSEF_rel"])),
            "HEIGHT_rel": np.full(n, float(canopy_evolution["HEIGHT_rel"])),
            "LAI_rel": lai_rel,
            "SAI_rel": np.full(n, float(canopy_evolution["SAI_rel"]))}


def _linear_interp_flat(knots, values, x):
    """Interpolations.linear_interpolation(knots, values; extrapolation_bc=Flat()): value is
    v[i]*(1-w) + v[i+1]*w with w = (x-k[i])/(k[i+1]-k[i])."""
    out = np.empty_like(x)
    for j, xv in enumerate(x):
        if xv <= knots[0]:
            out[j] = values[0]
        elif xv >= knots[-1]:
            out[j] = values[-1]
        else:
            i = int(np.searchsorted(knots, xv, side="right")) - 1
            w = (xv - knots[i]) / (knots[i + 1] - knots[i])
            out[j] = values[i] * (1 - w) + values[i + 1] * w
    return out


def build_forcing_table(spac: SPAC) -> np.ndarray:
    """[n_days_available + 1, LWFB_F] table; row d = what the reference evaluates at t = d."""
    F = lo.F
    meteo = spac.forcing["meteo"]
    nd = len(meteo["days"])
    nrows = nd + 1
    tab = np.zeros((nrows, lo.NF))
    dup = lambda a: np.concatenate([a, a[-1:]])        # last day duplicated (:711)
    tab[:, F.GLOBRAD] = dup(meteo["GLOBRAD"])
    tab[:, F.TMAX] = dup(meteo["TMAX"])
    tab[:, F.TMIN] = dup(meteo["TMIN"])
    tab[:, F.VAPPRES] = dup(meteo["VAPPRES"])
    tab[:, F.WIND] = dup(meteo["WIND"])
    tab[:, F.PREC] = dup(meteo["PRECIN"])
    day0 = meteo["days"][0]
    days_ext = day0 + np.arange(nrows, dtype=float)
    can = canopy_relative(spac.pars["canopy_evolution"], days_ext, spac.reference_date)
    for col, key in ((F.DENSEF_REL, "DENSEF_rel"), (F.HEIGHT_REL, "HEIGHT_rel"),
                     (F.LAI_REL, "LAI_rel"), (F.SAI_REL, "SAI_rel")):
        v = np.asarray(can[key], float) / 100
        if len(v) == nd:                                 # table from meteoveg.csv: 0 beyond the end (:850-855)
            v = np.concatenate([v, [0.0]])
        tab[:, col] = v
    storm = spac.forcing["storm_durations"]
    for r in range(nrows):
        t = float(days_ext[r])
        tab[r, F.DOY] = doy_of(spac.reference_date, t)
        mon = month_of(spac.reference_date, t)
        tab[r, F.IHD] = math.floor((storm[mon - 1] + 0.1) / 2.0)
        if spac.forcing["meteo_iso"] is not None:
            tab[r, F.D18O] = iso_prec_at(spac.forcing["meteo_iso"], t, "delta18O_permil")
            tab[r, F.D2H] = iso_prec_at(spac.forcing["meteo_iso"], t, "delta2H_permil")
        else:
            tab[r, F.D18O] = tab[r, F.D2H] = float("nan")
    if spac.forcing["meteo_iso"] is not None:
        frac = np.mod(spac.forcing["meteo_iso"]["days"] + 0.5, 1.0)
        if not np.all(frac == 0.5):
            raise ValueError("isotope samples must be dated at midnight (knots at integer day + 0.5) "
                             "for the per-day forcing table")
    return tab


# ----------------------------------------------------------------------------- parameters

def EQUIVSLP(LAT, SLOPE, ASPECT):
    L1 = math.asin(math.cos(SLOPE) * math.sin(LAT) + math.sin(SLOPE) * math.cos(LAT) * math.cos(ASPECT))
    D1 = math.cos(SLOPE) * math.cos(LAT) - math.sin(SLOPE) * math.sin(LAT) * math.cos(ASPECT)
    if D1 == 0:
        D1 = 0.0000000001
    L2 = math.atan(math.sin(SLOPE) * math.sin(ASPECT) / D1)
    if D1 < 0:
        L2 = L2 + P_PI
    return L1, L2


def INFPAR(INFEXP, ILAYER, THICK):
    N = len(THICK)
    inf = np.zeros(N)
    if INFEXP <= 0 or ILAYER <= 0:
        inf[0] = 1.0
    else:
        THICKT = float(np.sum(THICK[:ILAYER]))
        THICKA = np.cumsum(THICK[:ILAYER])
        inf[0] = (THICKA[0] / THICKT) ** INFEXP - (0 / THICKT) ** INFEXP
        for i in range(1, N):
            if i + 1 <= ILAYER:
                inf[i] = (THICKA[i] / THICKT) ** INFEXP - (THICKA[i - 1] / THICKT) ** INFEXP
    return inf


def flatten_parameters(spac: SPAC, df, dz, IDEPTH_idx, QDEPTH_idx):
    """(scal[PS], lay[PL,N], p_soil) for one member (reference define_LWFB90_p)."""
    S, L = lo.S, lo.L
    par = spac.pars["params"]
    so = spac.solver_options
    N = len(df["Upper_m"])
    THICK = 1000 * np.asarray(dz, float)
    assert not np.any(np.isclose(THICK, 0.0))
    shp = df["shp"]
    p_soil = SoilPar(
        THICK=THICK,
        STONEF=[h.p_STONEF for h in shp], THSAT=[h.p_THSAT for h in shp],
        KSAT=[h.p_KSAT for h in shp], alpha=[h.p_MvGalpha for h in shp],
        n=[h.p_MvGn for h in shp], m=[h.p_MvGm for h in shp], l=[h.p_MvGl for h in shp],
        thr=[h.p_thr for h in shp], Kfc=2.0)

    scal = np.zeros(lo.NPS)
    scal[S.DPSIMAX], scal[S.DSWMAX], scal[S.DTIMAX] = so["DPSIMAX"], so["DSWMAX"], so["DTIMAX"]
    scal[S.VXYLEM] = par["VXYLEM_mm"]
    scal[S.DISPERSIVITY] = par["DISPERSIVITY_mm"] / 1000
    LAT = par["LAT_DEG"] / 57.296
    ASPECT = par["ASPECT_DEG"] / 57.296
    ESLOPE = par["ESLOPE_DEG"] / 57.296
    L1, L2 = EQUIVSLP(LAT, ESLOPE, ASPECT)
    scal[S.LAT], scal[S.ESLOPE], scal[S.L1], scal[S.L2] = LAT, ESLOPE, L1, L2
    direct = {
        "MELFAC": "MELFAC", "RSTEMP": "RSTEMP", "FETCH": "FETCH", "WNDRAT": "WNDRAT", "Z0G": "Z0G",
        "Z0S": "Z0S", "Z0W": "Z0W", "ZMINH": "ZMINH", "ZW": "ZW", "HS": "HS", "HR": "HR",
        "CZS": "CZS", "CZR": "CZR", "C1": "C1", "C2": "C2", "C3": "C3", "SNODEN": "SNODEN",
        "CCFAC": "CCFAC", "GRDMLT": "GRDMLT", "LAIMLT": "LAIMLT", "SAIMLT": "SAIMLT",
        "MAXLQF": "MAXLQF", "CR": "CR", "LWIDTH": "LWIDTH", "RHOTP": "RHOTP", "LPC": "LPC",
        "KSNVP": "KSNVP", "ALBSN": "ALBSN", "ALB": "ALB", "NN": "NN", "RM": "RM",
        "CINTRL": "CINTRL", "CINTRS": "CINTRS", "CINTSL": "CINTSL", "CINTSS": "CINTSS",
        "FRINTL": "FRINTLAI", "FRINTS": "FRINTSAI", "FSINTL": "FSINTLAI", "FSINTS": "FSINTSAI",
        "MXKPL": "MXKPL", "FXYLEM": "FXYLEM", "GLMAX": "GLMAX", "GLMIN": "GLMIN", "TH": "TH",
        "T1": "T1", "T2": "T2", "TL": "TL", "PSICR": "PSICR", "CVPD": "CVPD", "R5": "R5",
        "RTRAD": "RTRAD", "MXRTLN": "MXRTLN", "RSSA": "RSSA", "RSSB": "RSSB", "IMPERV": "IMPERV",
        "BYPAR": "BYPAR", "DRAIN": "DRAIN", "DSLOPE": "DSLOPE", "GSC": "GSC", "GSP": "GSP",
        "LENGTH_SLOPE": "LENGTH_SLOPE", "QFFC": "QFFC", "QFPAR": "QFPAR",
        "MAXLAI": "MAXLAI", "SAI_BASELINE": "SAI_baseline_", "DENSEF_BASELINE": "DENSEF_baseline_",
        "HEIGHT_BASELINE": "HEIGHT_baseline_m", "AGE0": "AGE_baseline_yrs", "RGROPER": "RGROPER",
    }
    for sname, pname in direct.items():
        scal[getattr(S, sname)] = float(par[pname])
    scal[S.NOOUTF] = 1.0 if int(par["NOOUTF"]) == 1 else 0.0
    scal[S.QLAYER] = float(QDEPTH_idx)
    scal[S.SWATQX] = float(np.sum(p_soil.p_SWATMAX[:QDEPTH_idx]))
    scal[S.SWATQF] = float(np.sum(p_soil.p_THETAF[:QDEPTH_idx] * p_soil.p_THICK[:QDEPTH_idx]
                                  * (1 - p_soil.p_STONEF[:QDEPTH_idx])))
    scal[S.RL0] = par["INITRLEN"] / par["INITRDEP"]

    prof, t_ini = root_growth_tini(THICK, df["Rootden_"], par["INITRDEP"], par["RGRORATE"], par["RGROPER"])
    # RELDEN(t) is constant if every rooted layer is past its growth period for the whole record
    # (reference HammelKennel_lateral_rootgrowth, src/func_DiffEq_definition_p.jl:606-628)
    age0 = float(par["AGE_baseline_yrs"])
    finite = t_ini[t_ini < 1.0e19]
    if par["RGROPER"] <= 0:
        const = True
    else:
        const = bool(len(finite) == 0 or age0 > np.max(finite) + par["RGROPER"])
    scal[S.RELDEN_CONST] = 1.0 if const else 0.0

    lay = np.zeros((lo.NPL, N))
    lay[L.THICK] = p_soil.p_THICK
    lay[L.STONEF] = p_soil.p_STONEF
    lay[L.THSAT] = p_soil.p_THSAT
    lay[L.KSAT] = p_soil.p_KSAT
    lay[L.ALPHA] = p_soil.p_MvGalpha
    lay[L.NPAR] = p_soil.p_MvGn
    lay[L.MPAR] = p_soil.p_MvGm
    lay[L.TORT] = p_soil.p_MvGl
    lay[L.THR] = p_soil.p_thr
    lay[L.PSIF] = p_soil.p_PSIF
    lay[L.THETAF] = p_soil.p_THETAF
    lay[L.PSIG] = p_soil.p_PSIG
    lay[L.SWATMAX] = p_soil.p_SWATMAX
    lay[L.WETF] = p_soil.p_WETF
    lay[L.INFRAC] = INFPAR(par["INFEXP"], IDEPTH_idx, p_soil.p_THICK)
    lay[L.RELDEN] = prof
    lay[L.TINI] = t_ini
    lay[L.DZ, :N - 1] = p_soil.dz_m
    return scal, lay, p_soil


def initial_state(spac: SPAC, df, p_soil: SoilPar) -> np.ndarray:
    """u0 in the reference's state order (reference init_LWFB90_u0!, :84-139)."""
    iso = bool(spac.solver_options["simulate_isotopes"])
    N = p_soil.NLAYER
    ix = lo.StateIndex(N, iso)
    u0 = np.zeros(ix.L)
    psim0 = np.asarray(df["uAux_PSIM_init_kPa"], float)
    W0 = FWETNES(psim0, p_soil.p_MvGalpha, p_soil.p_MvGn, p_soil.p_MvGm)
    swati0 = FTheta_MvG(W0, p_soil.p_THSAT, p_soil.p_thr) * p_soil.p_SWATMAX / p_soil.p_THSAT
    ic = spac.pars["IC_scalar"]
    u0[ix.SWATI_mm] = swati0
    for name in ("GWAT", "INTS", "INTR", "SNOW"):
        col = ic[f"u_{name}_init_mm"]
        u0[getattr(ix, name + "_mm")] = col[0]
        if iso:
            u0[getattr(ix, name + "_d18O")] = col[1]
            u0[getattr(ix, name + "_d2H")] = col[2]
    u0[ix.CC] = ic["u_CC_init_MJ_per_m2"][0]
    u0[ix.SNOWLQ] = ic["u_SNOWLQ_init_mm"][0]
    u0[ix.RWU_mm] = 0.0
    u0[ix.XYLEM_mm] = 5.0
    if iso:
        d18 = np.asarray(df["u_delta18O_init_permil"], float)
        d2 = np.asarray(df["u_delta2H_init_permil"], float)
        u0[ix.SWATI_d18O] = d18
        u0[ix.SWATI_d2H] = d2
        u0[ix.RWU_d18O], u0[ix.RWU_d2H] = d18[0], d2[0]
        u0[ix.XYLEM_d18O], u0[ix.XYLEM_d2H] = d18[0], d2[0]
        u0[ix.TRANI_d18O] = d18
        u0[ix.TRANI_d2H] = d2
    new_swat = float(np.sum(swati0))
    u0[ix.acc("StorageSWAT")] = new_swat
    u0[ix.acc("StorageWATER")] = (u0[ix.INTR_mm] + u0[ix.INTS_mm] + u0[ix.SNOW_mm] + new_swat + u0[ix.GWAT_mm])
    return u0


# ----------------------------------------------------------------------------- setup / remake

def setup(parametrizedSPAC: SPAC, *, requested_tspan=None, soil_output_depths_m=(), eps=0.050,
          ε=None) -> DiscretizedSPAC:
    """Same arguments and meaning as the reference's setup (src/LWFBrook90.jl:365-370)."""
    if ε is not None:
        eps = ε
    soil_output_depths_m = list(soil_output_depths_m)
    assert all(d < 0 for d in soil_output_depths_m)
    spac = parametrizedSPAC.copy()
    if requested_tspan is not None:
        if isinstance(requested_tspan[0], _dt.datetime):
            requested_tspan = tuple((d - spac.reference_date).total_seconds() / 86400.0 for d in requested_tspan)
        lo_d = float(spac.forcing["meteo"]["days"].min())
        hi_d = float(spac.forcing["meteo"]["days"].max()) + 1.0
        if requested_tspan[0] < lo_d or requested_tspan[1] > hi_d:
            raise ValueError(f"Requested simulation tspan {tuple(requested_tspan)} goes beyond input forcing data: ({lo_d}, {hi_d})")
    par = spac.pars["params"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df_ref, IDEPTH_idx, QDEPTH_idx = refine_soil_discretization(
            spac.soil_discretization["df"], spac.pars["soil_horizons"], soil_output_depths_m,
            par["IDEPTH_m"], par["QDEPTH_m"], eps=eps)
    dz_ref = df_ref["Upper_m"] - df_ref["Lower_m"]
    if isinstance(spac.pars["root_distribution"], dict):
        overwrite_rootden(df_ref, spac.pars["root_distribution"], dz_ref)
        overwrite_IC(df_ref, spac.pars["IC_soil"], spac.solver_options["simulate_isotopes"])
    df_final = map_soil_horizons_to_discretization(spac.pars["soil_horizons"], df_ref)
    spac.soil_discretization = {"dz": dz_ref, "df": df_final}

    scal, lay, p_soil = flatten_parameters(spac, df_final, dz_ref, IDEPTH_idx, QDEPTH_idx)
    # canopy sanity checks of interpolate_aboveground_veg (:841-849)
    assert par["HEIGHT_baseline_m"] * 1.0 > 0.01
    u0 = initial_state(spac, df_final, p_soil)
    forcing = build_forcing_table(spac)

    tspan = spac.tspan if requested_tspan is None else (float(requested_tspan[0]), float(requested_tspan[1]))
    if requested_tspan is not None:
        spac.tspan = tspan
    if tspan[0] != math.floor(tspan[0]) or tspan[1] != math.floor(tspan[1]):
        raise ValueError("the batched integrator needs whole-day tspan bounds (saveat = t0:1:t1)")
    return DiscretizedSPAC(parametrizedSPAC=spac, N=p_soil.NLAYER,
                           iso=bool(spac.solver_options["simulate_isotopes"]), tspan=tspan,
                           scal=scal, lay=lay, u0=u0, forcing=forcing, p_soil=p_soil,
                           ILAYER=IDEPTH_idx, QLAYER=QDEPTH_idx)


_SHP_NAMES = {"ths_": "p_THSAT", "thr_": "p_thr", "Ksat_mmday": "p_KSAT", "alpha_per_m": "p_MvGalpha",
              "npar_": "p_MvGn", "tort_": "p_MvGl", "gravel_volFrac": "p_STONEF"}
_LAI_NAMES = ("DOY_Bstart", "Bduration", "DOY_Cstart", "Cduration", "LAI_perc_BtoC", "LAI_perc_CtoB")


def remakeSPAC(model, *, requested_tspan=None, soil_output_depths_m=(), **kwargs) -> DiscretizedSPAC:
    """Copy of a SPAC / DiscretizedSPAC with modified parameters, discretised again
    (reference remakeSPAC, src/LWFBrook90.jl:214-345).  kwargs: soil_horizons, LAI_rel, params,
    root_distribution, IC_soil, IC_scalar — each a dict."""
    base = model.parametrizedSPAC if isinstance(model, DiscretizedSPAC) else model
    spac = base.copy()
    for key, val in kwargs.items():
        if not isinstance(val, dict):
            raise AssertionError(f"Argument {key} is not a dict (NamedTuple in the reference)")
        if key == "soil_horizons":
            hz = spac.pars["soil_horizons"]
            nh = len(hz.shp)
            for k, v in val.items():
                assert k in _SHP_NAMES, f"Unclear how to remake '{k}' provided to soil_horizons."
                attr = _SHP_NAMES[k]
                if np.ndim(v) == 0:
                    ratio = v / getattr(hz.shp[0], attr)
                    target = [ratio * getattr(h, attr) for h in hz.shp]
                else:
                    assert len(v) == nh
                    target = list(v)
                for h, tv in zip(hz.shp, target):
                    setattr(h, attr, float(tv))
        elif key == "LAI_rel":
            ce = spac.pars["canopy_evolution"]
            assert not ce.get("_is_table", False), \
                "LAI evolution can only be modified if the SPAC was generated with a parametrized variant of LAI."
            for k in val:
                assert k in _LAI_NAMES, f"Unclear how to remake '{k}' provided to LAI."
            ce["LAI_rel"] = {**ce["LAI_rel"], **val}
        elif key == "params":
            for k in val:
                assert k in spac.pars["params"], f"Unclear how to remake '{k}' provided to params."
            spac.pars["params"] = {**spac.pars["params"], **val}
        elif key == "root_distribution":
            for k in val:
                assert k in ("beta", "z_rootMax_m")
            old = spac.pars["root_distribution"]
            old = old if isinstance(old, dict) else {}
            spac.pars["root_distribution"] = {**old, **val}
        elif key == "IC_soil":
            old = spac.pars["IC_soil"]
            for k in val:
                assert isinstance(old, dict) and k in old
            spac.pars["IC_soil"] = {**old, **val}
        elif key == "IC_scalar":
            spac.pars["IC_scalar"] = copy.deepcopy(val)
        else:
            raise ValueError(f"Unknown argument provided to remake(): {key}")
    return setup(spac, requested_tspan=requested_tspan, soil_output_depths_m=soil_output_depths_m)
