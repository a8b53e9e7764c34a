"""Host-side twin of the reference's input layer: ``loadSPAC`` and its CSV readers.

Follows (does not copy) src/func_read_inputData.jl of LWFBrook90.jl v0.9.9:
  loadSPAC                      :74-299
  read_path_meteoveg            :716-816
  read_path_meteoiso            :818-944
  read_path_initial_conditions  :946-969
  read_path_param               :978-1059
  read_path_storm_durations     :1061-1087
  read_path_soil_horizons       :1089-1164
  read_path_soil_discretization :1166-1207
  p_DOY / p_MONTHN              :651-661

Needed here only because Julia is not available in the build container; with Julia present the
reference's own loadSPAC/setup feed the C ABI directly.
"""
from __future__ import annotations

import copy
import csv
import datetime as _dt
import math
import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np

PARAM_KEYS = [
    "DISPERSIVITY_mm", "VXYLEM_mm", "LAT_DEG", "ESLOPE_DEG", "ASPECT_DEG", "ALB", "ALBSN",
    "C1", "C2", "C3", "WNDRAT", "FETCH", "Z0W", "ZW", "MAXLAI", "DENSEF_baseline_",
    "SAI_baseline_", "AGE_baseline_yrs", "HEIGHT_baseline_m", "LWIDTH", "Z0G", "Z0S", "LPC",
    "CZS", "CZR", "HS", "HR", "ZMINH", "RHOTP", "NN", "FRINTLAI", "FSINTLAI", "FRINTSAI",
    "FSINTSAI", "CINTRL", "CINTRS", "CINTSL", "CINTSS", "RSTEMP", "MELFAC", "CCFAC", "LAIMLT",
    "SAIMLT", "GRDMLT", "MAXLQF", "KSNVP", "SNODEN", "GLMAX", "GLMIN", "CR", "RM", "R5", "CVPD",
    "TL", "T1", "T2", "TH", "MXKPL", "MXRTLN", "INITRLEN", "INITRDEP", "RGRORATE", "RGROPER",
    "FXYLEM", "PSICR", "RTRAD", "NOOUTF", "IDEPTH_m", "QDEPTH_m", "RSSA", "RSSB", "INFEXP",
    "BYPAR", "QFPAR", "QFFC", "IMPERV", "DSLOPE", "LENGTH_SLOPE", "DRAIN", "GSC", "GSP",
    "DTIMAX", "DSWMAX", "DPSIMAX",
]
SOLVER_KEYS = ("DTIMAX", "DSWMAX", "DPSIMAX")

SHP_COLS = ("ths_volFrac", "thr_volFrac", "alpha_perMeter", "npar_", "ksat_mmDay", "tort_",
            "gravel_volFrac")

IC_SCALAR_COLS = ("u_GWAT_init_mm", "u_INTS_init_mm", "u_INTR_init_mm", "u_SNOW_init_mm",
                  "u_CC_init_MJ_per_m2", "u_SNOWLQ_init_mm")


def _parse_date(s: str) -> _dt.datetime:
    s = s.strip()
    for fmt in ("%Y-%m-%d", "%Y-%m-%dT%H:%M:%S", "%Y-%m-%dT%H:%M:%S.%f", "%Y-%m-%d %H:%M:%S"):
        try:
            return _dt.datetime.strptime(s, fmt)
        except ValueError:
            pass
    raise ValueError(f"cannot parse date {s!r}")


def _rows(path: str) -> List[List[str]]:
    with open(path, newline="") as fh:
        return [r for r in csv.reader(fh) if len(r) > 0]


@dataclass
class MvGHorizon:
    """One soil horizon with Mualem-van Genuchten hydraulic parameters
    (reference: MualemVanGenuchtenSHP, src/module_KPT.jl:97-129)."""
    p_THSAT: float
    p_thr: float
    p_MvGalpha: float
    p_MvGn: float
    p_KSAT: float
    p_MvGl: float
    p_STONEF: float

    @property
    def p_MvGm(self) -> float:
        return 1.0 - 1.0 / self.p_MvGn


@dataclass
class SoilHorizons:
    HorizonNr: List[int]
    Upper_m: List[float]
    Lower_m: List[float]
    shp: List[MvGHorizon]


@dataclass
class SPAC:
    """Parametrised soil-plant-atmosphere continuum (reference: struct SPAC, src/LWFBrook90.jl:65-82)."""
    reference_date: _dt.datetime
    tspan: tuple
    solver_options: Dict[str, Any]
    soil_discretization: Dict[str, Any]   # {"dz": ndarray, "df": dict of column arrays}
    forcing: Dict[str, Any]               # meteo (dict of daily arrays), meteo_iso, storm_durations
    pars: Dict[str, Any]                  # params, root_distribution, IC_scalar, IC_soil, canopy_evolution, soil_horizons

    def copy(self) -> "SPAC":
        return copy.deepcopy(self)


# ----------------------------------------------------------------------------- readers

def read_path_param(path: str):
    """key,value file with '###' comment rows (reference :978-1059)."""
    vals: Dict[str, float] = {}
    for r in _rows(path):
        key = r[0].strip()
        if key.startswith("###") or key == "param_id":
            continue
        vals[key] = float(r[1])
    missing = [k for k in PARAM_KEYS if k not in vals]
    extra = [k for k in vals if k not in PARAM_KEYS]
    if missing or extra:
        raise ValueError(f"{path}: missing keys {missing}, unexpected keys {extra}")
    vals["FXYLEM"] = min(vals["FXYLEM"], 0.990)
    vals["INITRLEN"] = max(vals["INITRLEN"], 0.010)
    vals["INITRDEP"] = max(vals["INITRDEP"], 0.010)
    vals["BYPAR"] = int(round(vals["BYPAR"]))
    vals["NOOUTF"] = int(round(vals["NOOUTF"]))
    solver = {k: vals[k] for k in SOLVER_KEYS}
    params = {k: vals[k] for k in PARAM_KEYS if k not in SOLVER_KEYS}
    return params, solver


def read_path_meteoveg(path: str):
    """Daily forcing, 7 or 11 columns, units row on line 2 (reference :716-816)."""
    rows = _rows(path)
    header = [h.strip() for h in rows[0]]
    base = ["dates", "globrad_MJDayM2", "tmax_degC", "tmin_degC", "vappres_kPa", "windspeed_ms",
            "prec_mmDay"]
    veg = ["densef_percent", "height_percent", "lai_percent", "sai_percent"]
    if header not in (base, base + veg):
        raise ValueError(f"{path}: unexpected columns {header}")
    data = rows[2:]
    dates = [_parse_date(r[0]) for r in data]
    for a, b in zip(dates[:-1], dates[1:]):
        if (b - a) != _dt.timedelta(days=1):
            raise ValueError(f"{path}: gaps in the forcing file")
    ref = dates[0]
    out = {"days": np.array([(d - ref).total_seconds() / 86400.0 for d in dates])}
    names = ["GLOBRAD", "TMAX", "TMIN", "VAPPRES", "WIND", "PRECIN"]
    if len(header) == 11:
        names += ["DENSEF_rel", "HEIGHT_rel", "LAI_rel", "SAI_rel"]
    for j, n in enumerate(names):
        out[n] = np.array([float(r[j + 1]) for r in data])
    return out, ref


def read_path_meteoiso(path: str, meteo, ref: _dt.datetime):
    """Isotopic signature of precipitation (reference :818-944).  Both the plain format and the
    Piso.AI format are handled; the first sample is overwritten by the second (:910-911)."""
    with open(path) as fh:
        first = fh.readline()
    if first.startswith("#Data produced with Piso.AI"):
        rows = _rows(path)
        hdr = [h.strip() for h in rows[3]]
        data = rows[4:]
        i_d, i_o, i_h = hdr.index("Date"), hdr.index("d18O.Piso.AI"), hdr.index("d2H.Piso.AI")

        def ceil_month(d):
            if d.day == 1 and d.time() == _dt.time(0):
                return d
            y, m = (d.year + 1, 1) if d.month == 12 else (d.year, d.month + 1)
            return _dt.datetime(y, m, 1)
        recs = []
        for r in data:
            if r[i_o] in ("", "NA") or r[i_h] in ("", "NA"):
                raise ValueError(f"{path}: missing isotope value")
            recs.append((_parse_date(r[i_d]), float(r[i_o]), float(r[i_h])))
        d0 = recs[0][0]
        firstrec = (_dt.datetime(d0.year, d0.month, 1), recs[0][1], recs[0][2])
        recs = [(ceil_month(d), o, h) for d, o, h in recs] + [firstrec]
        recs.sort(key=lambda x: x[0])
        recs[-1] = (recs[-1][0] - _dt.timedelta(days=1), recs[-1][1], recs[-1][2])
        dates = [r[0] for r in recs]
        d18 = [r[1] for r in recs]
        d2 = [r[2] for r in recs]
    else:
        rows = _rows(path)
        hdr = [h.strip() for h in rows[0]]
        if hdr != ["dates", "delta18O_permil", "delta2H_permil"]:
            raise ValueError(f"{path}: unexpected columns {hdr}")
        data = rows[2:]
        dates = [_parse_date(r[0]) for r in data]

        def num(s):
            s = s.strip()
            return float("nan") if s in ("", "NA") else float(s)
        d18 = [num(r[1]) for r in data]
        d2 = [num(r[2]) for r in data]
        d18[0] = d18[1]
        d2[0] = d2[1]
        if any(math.isnan(x) for x in d18 + d2):
            raise ValueError(f"{path}: missing isotope values")
    days = np.array([(d - ref).total_seconds() / 86400.0 for d in dates])
    return {"days": days, "delta18O_permil": np.array(d18), "delta2H_permil": np.array(d2)}


def read_path_storm_durations(path: str):
    vals = []
    months = []
    for r in _rows(path):
        k = r[0].strip()
        if k.startswith("###") or k == "month":
            continue
        months.append(k)
        vals.append(float(r[1]))
    expected = ["January", "Februrary", "March", "April", "May", "June", "July", "August",
                "September", "October", "November", "December"]
    if months != expected:
        raise ValueError(f"{path}: months not in the expected order")
    return np.array(vals)


def read_path_initial_conditions(path: str):
    """6 rows x (amount, d18O, d2H) (reference :946-969). Returns dict col -> array[3]."""
    out = {}
    for r in _rows(path):
        k = r[0].strip()
        if k.startswith("###") or k == "param_id":
            continue

        def num(s):
            s = s.strip()
            return float("nan") if s in ("", "NA") else float(s)
        out[k] = np.array([num(r[1]), num(r[2]), num(r[3])])
    if sorted(out) != sorted(IC_SCALAR_COLS):
        raise ValueError(f"{path}: unexpected rows {sorted(out)}")
    out["u_CC_init_MJ_per_m2"][1:] = -9999.99
    out["u_SNOWLQ_init_mm"][1:] = -9999.99
    return out


def read_path_soil_horizons(path: str) -> SoilHorizons:
    rows = _rows(path)
    hdr = [h.strip() for h in rows[0]]
    mvg = ["HorizonNr", "Upper_m", "Lower_m"] + list(SHP_COLS)
    if hdr != mvg:
        raise ValueError(f"{path}: only Mualem-van Genuchten horizon files are supported "
                         f"(Clapp-Hornberger is unreachable from the reference's RHS), got {hdr}")
    nr, up, lo, shp = [], [], [], []
    for r in rows[2:]:
        v = [float(x) for x in r]
        nr.append(int(v[0])); up.append(v[1]); lo.append(v[2])
        shp.append(MvGHorizon(p_THSAT=v[3], p_thr=v[4], p_MvGalpha=v[5], p_MvGn=v[6],
                              p_KSAT=v[7], p_MvGl=v[8], p_STONEF=v[9]))
    if lo[:-1] != up[1:]:
        raise ValueError(f"{path}: overlapping horizons")
    if abs(up[0]) > 1e-12:
        raise ValueError(f"{path}: first horizon must start at 0")
    return SoilHorizons(nr, up, lo, shp)


def read_path_soil_discretization(path: str):
    rows = _rows(path)
    hdr = [h.strip() for h in rows[0]]
    need = ["Upper_m", "Lower_m", "Rootden_", "uAux_PSIM_init_kPa"]
    for n in need:
        if n not in hdr:
            raise ValueError(f"{path}: column {n} missing")
    cols = {h: [] for h in hdr}
    for r in rows[2:]:
        for h, x in zip(hdr, r):
            x = x.strip()
            cols[h].append(float("nan") if x in ("", "NA") else float(x))
    df = {h: np.array(v) for h, v in cols.items()}
    if not np.array_equal(df["Lower_m"][:-1], df["Upper_m"][1:]):
        raise ValueError(f"{path}: overlapping layers")
    for c in ("u_delta18O_init_permil", "u_delta2H_init_permil"):
        if c not in df:
            df[c] = np.full(len(df["Upper_m"]), -9999.0)
    if np.isnan(df["u_delta18O_init_permil"]).any() or np.isnan(df["u_delta2H_init_permil"]).any():
        df["u_delta18O_init_permil"] = np.full(len(df["Upper_m"]), -9999.0)
        df["u_delta2H_init_permil"] = np.full(len(df["Upper_m"]), -9999.0)
    return df


# ----------------------------------------------------------------------------- calendar helpers

def doy_of(ref: _dt.datetime, t: float) -> int:
    """p_DOY (reference :651)"""
    return (ref + _dt.timedelta(days=math.floor(t))).timetuple().tm_yday


def month_of(ref: _dt.datetime, t: float) -> int:
    """p_MONTHN (reference :659)"""
    return (ref + _dt.timedelta(days=math.floor(t))).month


# ----------------------------------------------------------------------------- loadSPAC

def loadSPAC(folder: str, prefix: str, *, simulate_isotopes: bool = True,
             canopy_evolution="meteoveg.csv", Δz_thickness_m="soil_discretization.csv",
             root_distribution="soil_discretization.csv", IC_soil="soil_discretization.csv",
             IC_scalar="initial_conditions.csv", storm_durations_h="meteo_storm_durations.csv",
             dz_thickness_m=None) -> SPAC:
    """Same arguments and meaning as the reference's loadSPAC (src/func_read_inputData.jl:74-85).
    ``dz_thickness_m`` is an ASCII alias of ``Δz_thickness_m``."""
    from .discretize import extend_lowest_horizon, overwrite_IC, overwrite_rootden

    if dz_thickness_m is not None:
        Δz_thickness_m = dz_thickness_m
    p = lambda kind: os.path.join(folder, f"{prefix}_{kind}.csv")
    params, solver_opts = read_path_param(p("param"))
    solver_options = {"compute_intermediate_quantities": True,
                      "simulate_isotopes": bool(simulate_isotopes), **solver_opts}
    meteoveg, ref = read_path_meteoveg(p("meteoveg"))
    meteo_iso = read_path_meteoiso(p("meteoiso"), meteoveg, ref) if simulate_isotopes else None
    if isinstance(storm_durations_h, (list, tuple, np.ndarray)):
        if len(storm_durations_h) != 12:
            raise ValueError("storm_durations_h must have length 12")
        storm = np.array(storm_durations_h, dtype=float)
    elif storm_durations_h == "meteo_storm_durations.csv":
        storm = read_path_storm_durations(p("meteo_storm_durations"))
    else:
        raise ValueError(f"Unknown format for argument `meteo_storm_durations`: {storm_durations_h}")
    for k in ("GLOBRAD", "VAPPRES", "WIND", "PRECIN"):
        if not (meteoveg[k] >= 0).all():
            raise AssertionError(f"Error in forcing: {k} must be >= 0")

    if isinstance(canopy_evolution, str):
        if canopy_evolution != "meteoveg.csv":
            raise ValueError("canopy_evolution must be 'meteoveg.csv' or a dict")
        if "LAI_rel" not in meteoveg:
            raise ValueError("meteoveg.csv has no canopy columns; provide canopy_evolution=dict(...)")
        canopy = {k: meteoveg[k].copy() for k in ("days", "DENSEF_rel", "HEIGHT_rel", "LAI_rel", "SAI_rel")}
        canopy["_is_table"] = True
        assert (canopy["LAI_rel"] >= 0).all()
        chk = canopy
    else:
        canopy = copy.deepcopy(dict(canopy_evolution))
        if tuple(canopy.keys()) != ("DENSEF_rel", "HEIGHT_rel", "SAI_rel", "LAI_rel"):
            raise AssertionError("canopy_evolution keys must be (DENSEF_rel, HEIGHT_rel, SAI_rel, LAI_rel)")
        if tuple(canopy["LAI_rel"].keys()) != ("DOY_Bstart", "Bduration", "DOY_Cstart", "Cduration",
                                               "LAI_perc_BtoC", "LAI_perc_CtoB"):
            raise AssertionError("canopy_evolution.LAI_rel has unexpected keys")
        chk = canopy
    assert np.all(np.asarray(chk["DENSEF_rel"]) > 5), "DENSEF should not be set lower than 5%"
    assert np.all(np.asarray(chk["HEIGHT_rel"]) > 0)
    assert np.all(np.asarray(chk["SAI_rel"]) > 0)

    soil_horizons = read_path_soil_horizons(p("soil_horizons"))

    if isinstance(IC_scalar, dict):
        a, o, h = IC_scalar["amount"], IC_scalar["d18O"], IC_scalar["d2H"]
        nan = float("nan")
        ic = {
            "u_GWAT_init_mm": np.array([a["u_GWAT_init_mm"], o["u_GWAT_init_permil"], h["u_GWAT_init_permil"]], float),
            "u_INTS_init_mm": np.array([a["u_INTS_init_mm"], o["u_INTS_init_permil"], h["u_INTS_init_permil"]], float),
            "u_INTR_init_mm": np.array([a["u_INTR_init_mm"], o["u_INTR_init_permil"], h["u_INTR_init_permil"]], float),
            "u_SNOW_init_mm": np.array([a["u_SNOW_init_mm"], o["u_SNOW_init_permil"], h["u_SNOW_init_permil"]], float),
            "u_CC_init_MJ_per_m2": np.array([a["u_CC_init_MJ_per_m2"], nan, nan], float),
            "u_SNOWLQ_init_mm": np.array([a["u_SNOWLQ_init_mm"], nan, nan], float),
        }
    elif IC_scalar == "initial_conditions.csv":
        ic = read_path_initial_conditions(p("initial_conditions"))
    else:
        raise ValueError(f"Unknown format for argument `IC_scalar`: {IC_scalar}")

    if isinstance(Δz_thickness_m, (list, tuple, np.ndarray)):
        dz = np.array(Δz_thickness_m, dtype=float)
        interfaces = -np.concatenate([[0.0], np.cumsum(dz)])
        n = len(dz)
        df = {"Upper_m": interfaces[:-1].copy(), "Lower_m": interfaces[1:].copy(),
              "Rootden_": np.full(n, np.nan), "uAux_PSIM_init_kPa": np.full(n, np.nan),
              "u_delta18O_init_permil": np.full(n, np.nan), "u_delta2H_init_permil": np.full(n, np.nan)}
        if isinstance(root_distribution, str):
            raise ValueError("manual soil discretization requires a parametric root_distribution")
        if isinstance(IC_soil, str):
            raise ValueError("manual soil discretization requires parametric soil initial conditions")
        overwrite_rootden(df, root_distribution, dz)
        overwrite_IC(df, IC_soil, simulate_isotopes)
    elif Δz_thickness_m == "soil_discretization.csv":
        df = read_path_soil_discretization(p("soil_discretization"))
        dz = df["Upper_m"] - df["Lower_m"]
        if not isinstance(root_distribution, str):
            overwrite_rootden(df, root_distribution, dz)
        if not isinstance(IC_soil, str):
            overwrite_IC(df, IC_soil, simulate_isotopes)
    else:
        raise ValueError(f"Unknown format for argument `Δz_thickness_m`: {Δz_thickness_m}")

    ext_horizons = extend_lowest_horizon(soil_horizons, df)
    ndays = len(meteoveg["days"])
    tspan = (float(meteoveg["days"].min()), float(meteoveg["days"].max()) + 1.0)
    meteo = {k: meteoveg[k].copy() for k in ("days", "GLOBRAD", "TMAX", "TMIN", "VAPPRES", "WIND", "PRECIN")}
    return SPAC(
        reference_date=ref, tspan=tspan, solver_options=solver_options,
        soil_discretization={"dz": np.array(dz, float), "df": df},
        forcing={"meteo": meteo, "meteo_iso": meteo_iso, "storm_durations": storm},
        pars={"params": params,
              "root_distribution": copy.deepcopy(root_distribution),
              "IC_scalar": ic,
              "IC_soil": copy.deepcopy(IC_soil),
              "canopy_evolution": canopy,
              "soil_horizons": ext_horizons})
