"""Host-side twin of the reference's soil-domain discretisation.

Follows src/func_discretize_soil_domain.jl of LWFBrook90.jl v0.9.9:
  extend_lowest_horizon               :70-96
  refine_soil_discretization          :111-243
  map_soil_horizons_to_discretization :250-306
  overwrite_IC! / overwrite_rootden!  :307-337
  Rootden_beta_                       :427-509
  HammelKennel_transient_root_density :550-645 (only the t_ini part; RELDEN(t) itself is
                                       evaluated on the device once a day)
"""
from __future__ import annotations

import copy
import warnings

import numpy as np

SOIL_COLS = ("Upper_m", "Lower_m", "Rootden_", "uAux_PSIM_init_kPa",
             "u_delta18O_init_permil", "u_delta2H_init_permil")


def _isapprox(a, b):
    # Julia's isapprox default: rtol = sqrt(eps), atol = 0
    return abs(a - b) <= np.sqrt(np.finfo(float).eps) * max(abs(a), abs(b))


def extend_lowest_horizon(soil_horizons, df):
    """reference :70-96"""
    ext = copy.deepcopy(soil_horizons)
    extent = (max(ext.Upper_m), min(ext.Lower_m))
    assert _isapprox(extent[0], df["Upper_m"][0]) or (extent[0] == 0 and df["Upper_m"][0] == 0), \
        "Spatial domain of soil discretization is not compatible with provided soil horizons"
    low = float(df["Lower_m"][-1])
    if _isapprox(extent[1], low):
        pass
    elif extent[1] < low:
        warnings.warn("Spatial domain soil discretization is smaller than the provided soil horizon information")
    else:
        warnings.warn("Lowest soil horizon will be extended downward to the soil discretization")
        ext.HorizonNr.append(ext.HorizonNr[-1] + 1)
        ext.Upper_m.append(ext.Lower_m[-1])
        ext.Lower_m.append(low)
        ext.shp.append(copy.deepcopy(ext.shp[-1]))
    return ext


def refine_soil_discretization(df, soil_horizons, soil_output_depths_m, IDEPTH_m, QDEPTH_m, eps=0.050):
    """Adds interfaces where a horizon begins, where IDEPTH/QDEPTH need one and where an output
    depth was requested (reference :111-243).  Returns (refined df, IDEPTH_idx, QDEPTH_idx)."""
    out_depths = np.asarray(soil_output_depths_m, dtype=float)
    if len(out_depths) > 1:
        assert np.all(np.abs(np.diff(out_depths)) >= eps), \
            f"Requested soil_output_depths_m must be further away from each other than {eps} m"
    interfaces_for_layers = sorted(np.concatenate([out_depths, out_depths - eps]).tolist(), reverse=True)
    interfaces_for_parameters = sorted(list(soil_horizons.Upper_m) + list(soil_horizons.Lower_m)
                                       + [-IDEPTH_m, -QDEPTH_m], reverse=True)
    allp = sorted(interfaces_for_layers + interfaces_for_parameters, reverse=True)
    uniq = []
    for x in allp:                       # unique() keeps first occurrences
        if x not in uniq:
            uniq.append(x)
    existing = np.concatenate([[0.0], df["Lower_m"]])
    needed = [x for x in uniq if not np.any(np.abs(x - existing) < eps * 0.5)]
    if len(needed) > 0:
        keep = [True] + list(np.abs(np.diff(needed)) >= eps * 0.9)
        needed = [x for x, k in zip(needed, keep) if k]

    cols = {k: list(v) for k, v in df.items()}
    n_of = lambda: len(cols["Upper_m"])
    added = []
    for itf in needed:
        for i in range(n_of()):
            if itf > cols["Lower_m"][i]:
                pos = i + 1
                for k in cols:
                    cols[k].insert(pos, cols[k][pos - 1])   # properties from the cell above
                cols["Lower_m"][pos - 1] = itf
                cols["Upper_m"][pos] = itf
                added.append(itf)
                break
            elif i == n_of() - 1:
                if itf >= cols["Lower_m"][i] - eps:
                    pos = i + 1
                    for k in cols:
                        cols[k].append(cols[k][pos - 1])
                    cols["Upper_m"][pos] = cols["Lower_m"][pos - 1]
                    cols["Lower_m"][pos] = itf
                    added.append(itf)
                    break
    if added:
        warnings.warn(f"Adding interfaces to soil discretization at depths {added} (m)")
    ref = {}
    for k, v in cols.items():
        if k == "shp":
            ref[k] = list(v)
        else:
            ref[k] = np.array(v, dtype=float)
    lower = ref["Lower_m"]
    IDEPTH_idx = int(np.sum(-IDEPTH_m <= lower))
    QDEPTH_idx = int(np.sum(-QDEPTH_m <= lower))
    if (-IDEPTH_m < lower[-1]) or (-QDEPTH_m < lower[-1]):
        raise ValueError("QDEPTH_m or IDEPTH_m were defined deeper than the lowest simulation element.")
    return ref, IDEPTH_idx, QDEPTH_idx


def map_soil_horizons_to_discretization(soil_horizons, grid):
    """Attributes horizon parameters to computational cells (reference :250-306)."""
    n = len(grid["Lower_m"])
    hl = np.array(soil_horizons.Lower_m)
    which = []
    for i in range(n):
        hit = np.nonzero(grid["Lower_m"][i] >= hl)[0]
        if len(hit) == 0:
            if i == n - 1:
                idx = len(hl) - 1
            else:
                raise ValueError("soil horizons do not cover the discretization domain")
        else:
            idx = int(hit[0])
        which.append(idx)
    out = {k: np.array(grid[k], dtype=float).copy() for k in SOIL_COLS}
    out["HorizonNr"] = np.array([soil_horizons.HorizonNr[i] for i in which])
    out["shp"] = [soil_horizons.shp[i] for i in which]
    thick_m = out["Upper_m"] - out["Lower_m"]
    assert np.all(thick_m > 0.0), "All discretized cells must be bigger than zero"
    assert np.all(out["uAux_PSIM_init_kPa"] <= 0), "Initial matrix psi must be negative or zero"
    return out


def overwrite_IC(df, IC_soil, simulate_isotopes):
    keys = tuple(IC_soil.keys())
    assert keys in (("PSIM_init_kPa", "delta18O_init_permil", "delta2H_init_permil"), ("PSIM_init_kPa",))
    assert not (simulate_isotopes and keys == ("PSIM_init_kPa",))
    n = len(df["Upper_m"])
    df["uAux_PSIM_init_kPa"] = np.full(n, float(IC_soil["PSIM_init_kPa"]))
    if len(keys) == 3:
        df["u_delta18O_init_permil"] = np.full(n, float(IC_soil["delta18O_init_permil"]))
        df["u_delta2H_init_permil"] = np.full(n, float(IC_soil["delta2H_init_permil"]))


def overwrite_rootden(df, root_distribution, dz):
    keys = tuple(root_distribution.keys())
    assert keys in (("beta", "z_rootMax_m"), ("beta",))
    zmax = root_distribution.get("z_rootMax_m", None)
    df["Rootden_"] = Rootden_beta_(root_distribution["beta"], dz_m=dz, z_rootMax_m=zmax)


def Rootden_beta_(beta, *, dz_m=None, z_rootMax_m=None, z_Upper_m=0.0, Δz_m=None):
    """Relative root density per layer from the beta model of Gale & Grigal (1987)
    (reference :427-509)."""
    if Δz_m is not None:
        dz_m = Δz_m
    dz_m = np.asarray(dz_m, dtype=float)
    if z_rootMax_m is None:
        z_rootMax_m = -np.sum(dz_m)
    assert z_Upper_m <= 0 and dz_m.min() > 0 and z_rootMax_m < 0
    z_if = z_Upper_m + np.concatenate([[0.0], np.cumsum(-dz_m)])
    Y = lambda d_cm: 1.0 - beta ** d_cm
    d_if = -z_if * 100
    d_root = -z_rootMax_m * 100
    inc = np.diff(Y(d_if))
    d_up, d_lo = d_if[:-1], d_if[1:]
    inc[d_up >= d_root] = 0
    hit = np.nonzero(d_root <= d_lo)[0]
    if len(hit) > 0:
        i = int(hit[0])
        inc[i] = Y(d_root) - Y(d_up[i])
    per_cm = inc / (dz_m * 100)
    assert abs(beta - 1.0) > 0.000001, "Root distribution parameter beta is too close to 1.0"
    s = per_cm.sum()
    return per_cm if s == 0 else per_cm / s


def root_growth_tini(THICK_mm, final_rootden, INITRDEP, RGRORATE, RGROPER):
    """Stand age at which roots reach each layer, and the clipped final profile
    (reference HammelKennel_transient_root_density :550-612).  Returns (final_profile, t_ini)."""
    prof = np.array(final_rootden, dtype=float).copy()
    N = len(prof)
    gt = np.nonzero(prof > 1.0e-6)[0]
    if len(gt) > 0:
        i1, i2 = int(gt[0]), int(gt[-1])
        for i in range(i1, i2 + 1):
            prof[i] = max(prof[i], 1.01e-6)
    t_ini = np.full(N, 1.0e20)
    if RGROPER > 0:
        THICK = np.asarray(THICK_mm, dtype=float)
        z_center = -(np.cumsum(THICK) - THICK / 2) / 1000
        depfirst = z_center[0] - THICK[0] / 1000.0
        for i in range(N):
            z = depfirst - z_center[i]
            if prof[i] >= 1.0e-6:
                if z <= INITRDEP:
                    t_ini[i] = 0.0
                elif RGRORATE > 0:
                    t_ini[i] = (z - INITRDEP) / RGRORATE
    else:
        t_ini[:] = 0.0
    return prof, t_ini
