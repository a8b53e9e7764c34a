"""Synthetic site x parameter-set ensembles and batched `setup`.

The reference has no ensemble generator; its only mechanism is remakeSPAC -> setup per member
(src/LWFBrook90.jl:223-254), which does DataFrame work and a bisection per layer and would dominate
wall time at 10^5..10^6 members.  `make_ensemble` draws the perturbations with a counter-based RNG
(Philox4x32-10, key = seed, counter = member index: CPU and GPU sides can regenerate any member
independently) and applies them with a vectorised twin of remake_* + define_LWFB90_p +
KPT_SOILPAR_Mvg1d + init_LWFB90_u0! (SURVEY.md section 8(d) lists the knobs; section 8(f) N1).
Member 0 is always the unmodified site definition.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

from . import layout as lo
from .discretize import Rootden_beta_, root_growth_tini
from .kpt import FK_MvG, FPSIM_MvG, FTheta_MvG, RHOWG
from .setup import DiscretizedSPAC, INFPAR


# ----------------------------------------------------------------------------- Philox4x32-10
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(counter: np.ndarray, key) -> np.ndarray:
    """counter: uint32 [n,4]; key: (k0,k1). Returns uint32 [n,4] (Salmon et al. 2011)."""
    c = counter.astype(np.uint32).copy()
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c[:, 0].astype(np.uint64)
            p1 = _M1 * c[:, 2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
            k0 = np.uint32(k0 + _W0)
            k1 = np.uint32(k1 + _W1)
    return c


def uniforms(member_idx: np.ndarray, n_per_member: int, seed: int = 1234) -> np.ndarray:
    """[n_members, n_per_member] doubles in [0,1): 53-bit from two Philox words;
    counter = (member, draw block, 0, 0), key = (seed, 0)."""
    mi = np.asarray(member_idx, dtype=np.uint64)
    nblk = (n_per_member + 1) // 2
    out = np.empty((len(mi), nblk * 2))
    for b in range(nblk):
        ctr = np.stack([(mi & np.uint64(0xFFFFFFFF)).astype(np.uint32), (mi >> np.uint64(32)).astype(np.uint32),
                        np.full(len(mi), b, np.uint32), np.zeros(len(mi), np.uint32)], axis=1)
        r = philox4x32_10(ctr, (seed, 0)).astype(np.uint64)
        for j in range(2):
            hi, lw = r[:, 2 * j], r[:, 2 * j + 1]
            out[:, 2 * b + j] = ((hi >> np.uint64(5)) * np.float64(67108864.0) + (lw >> np.uint64(6))) / 9007199254740992.0
    return out[:, :n_per_member]


# ----------------------------------------------------------------------------- batched setup
def _bisect_wetf_vec(Kfc, KSAT, l, n):
    """Vectorised twin of kpt._bisect_wetf (same midpoint sequence).  Large batches are split along the member axis
    and bisected on a thread pool (numpy releases the GIL inside the ufuncs; every operation is elementwise, so the
    result does not depend on the split)."""
    M = KSAT.shape[0]
    nthr = min(os.cpu_count() or 1, 16)
    if KSAT.ndim < 1 or M < 4096 or nthr < 2:
        return _bisect_wetf_chunk(Kfc, KSAT, l, n)
    bounds = np.linspace(0, M, 4 * nthr + 1).astype(int)
    with ThreadPoolExecutor(nthr) as pool:
        parts = list(pool.map(lambda k: _bisect_wetf_chunk(Kfc, KSAT[bounds[k]:bounds[k + 1]], l[bounds[k]:bounds[k + 1]],
                                                           n[bounds[k]:bounds[k + 1]]), range(4 * nthr)))
    return np.concatenate(parts, axis=0)


def _bisect_wetf_chunk(Kfc, KSAT, l, n):
    a = np.zeros_like(KSAT)
    b = np.ones_like(KSAT)
    f = lambda x: -Kfc + FK_MvG(x, KSAT, l, n)
    fa, fb = f(a), f(b)
    if not np.all((fa < 0) & (fb > 0)):
        raise ValueError("Computed invalid p_WETF for some members")
    active = np.ones(a.shape, dtype=bool)
    for _ in range(200):
        mid = 0.5 * (a + b)
        active &= (mid > a) & (mid < b)
        if not active.any():
            break
        fm = f(mid)
        left = active & ((fm < 0) == (fa < 0))
        right = active & ~left
        a = np.where(left, mid, a); fa = np.where(left, fm, fa)
        b = np.where(right, mid, b); fb = np.where(right, fm, fb)
        exact = active & (fm == 0)
        a = np.where(exact, mid, a); b = np.where(exact, mid, b)
    return np.where(np.abs(fa) <= np.abs(fb), a, b)


@dataclass
class Ensemble:
    """Flat blocks of an ensemble, ready for the C ABI."""
    N: int
    iso: bool
    t0_day: int
    n_days: int
    forcing: np.ndarray          # [n_sites, n_rows, F]
    site_of_member: np.ndarray   # [M] int32
    scal: np.ndarray             # [M, PS]
    lay: np.ndarray              # [M, PL, N]
    u0: np.ndarray               # [M, L]

    @property
    def M(self) -> int:
        return self.scal.shape[0]

    def subset(self, idx) -> "Ensemble":
        idx = np.asarray(idx)
        return Ensemble(self.N, self.iso, self.t0_day, self.n_days, self.forcing, self.site_of_member[idx].copy(),
                        self.scal[idx].copy(), self.lay[idx].copy(), self.u0[idx].copy())


def batched_setup(base: DiscretizedSPAC, *, ths_f=None, ksat_f=None, alpha_f=None, nm1_f=None, gravel_f=None,
                  params=None, beta=None, psim_init=None) -> Ensemble:
    """Vectorised remakeSPAC+setup for M members that share grid, forcing and canopy phenology.

    ``*_f`` are per-member multiplicative factors of the TOP horizon value; all horizons are scaled
    proportionally, as remake_soil_horizons does for scalar arguments (src/LWFBrook90.jl:266-272).
    ``nm1_f`` scales (n-1) of the top horizon.  ``params`` maps LWFB_S names to per-member arrays.
    ``beta`` is the root-distribution parameter per member (keeps the base's z_rootMax_m);
    ``psim_init`` the uniform initial matric potential (kPa) per member.
    """
    S, L = lo.S, lo.L
    N, iso = base.N, base.iso
    some = [x for x in (ths_f, ksat_f, alpha_f, nm1_f, gravel_f, beta, psim_init) if x is not None]
    if params:
        some += list(params.values())
    M = len(some[0]) if some else 1
    scal = np.repeat(base.scal[None], M, axis=0)
    lay = np.repeat(base.lay[None], M, axis=0)
    one = np.ones(M)
    col = lambda f: (one if f is None else np.asarray(f, float))[:, None]

    THICK = lay[:, L.THICK]
    THSAT = np.minimum(lay[:, L.THSAT] * col(ths_f), 0.95) if ths_f is not None else lay[:, L.THSAT]
    KSAT = lay[:, L.KSAT] * col(ksat_f)
    ALPHA = lay[:, L.ALPHA] * col(alpha_f)
    if nm1_f is not None:
        n_top = base.lay[L.NPAR, 0]
        ratio = (1.0 + (n_top - 1.0) * np.asarray(nm1_f, float)) / n_top
        NPAR = lay[:, L.NPAR] * ratio[:, None]
    else:
        NPAR = lay[:, L.NPAR]
    MPAR = 1.0 - 1.0 / NPAR
    STONEF = np.minimum(lay[:, L.STONEF] * col(gravel_f), 0.95) if gravel_f is not None else lay[:, L.STONEF]
    TORT, THR = lay[:, L.TORT], lay[:, L.THR]

    soil_changed = any(x is not None for x in (ths_f, ksat_f, alpha_f, nm1_f, gravel_f))
    SWATMAX = THICK * THSAT * (1.0 - STONEF)
    if soil_changed:
        WETF = _bisect_wetf_vec(2.0, KSAT, TORT, NPAR)
        PSIF = FPSIM_MvG(WETF, ALPHA, NPAR, MPAR)
        THETAF = FTheta_MvG(WETF, THSAT, THR)
    else:
        WETF, PSIF, THETAF = lay[:, L.WETF], lay[:, L.PSIF], lay[:, L.THETAF]
    lay[:, L.THSAT], lay[:, L.KSAT], lay[:, L.ALPHA], lay[:, L.NPAR], lay[:, L.MPAR] = THSAT, KSAT, ALPHA, NPAR, MPAR
    lay[:, L.STONEF], lay[:, L.SWATMAX], lay[:, L.WETF], lay[:, L.PSIF], lay[:, L.THETAF] = STONEF, SWATMAX, WETF, PSIF, THETAF

    params = dict(params or {})
    infexp = params.pop("INFEXP", None)
    for name, v in params.items():
        scal[:, getattr(S, name)] = np.asarray(v, float)
    Q = int(base.scal[S.QLAYER])
    scal[:, S.SWATQX] = SWATMAX[:, :Q].sum(axis=1)
    scal[:, S.SWATQF] = (THETAF[:, :Q] * THICK[:, :Q] * (1 - STONEF[:, :Q])).sum(axis=1)
    if infexp is not None:   # WAT.INFPAR, src/module_WAT.jl:59-79, vectorised over members
        e = np.asarray(infexp, float)
        IL = int(base.ILAYER)
        inf = np.zeros((M, N))
        top_only = (e <= 0) | (IL <= 0)
        if IL > 0:
            THICKT = THICK[:, :IL].sum(axis=1)
            THICKA = np.cumsum(THICK[:, :IL], axis=1)
            frac = (THICKA / THICKT[:, None]) ** e[:, None]
            inf[:, 0] = frac[:, 0] - (0 / THICKT) ** e
            inf[:, 1:IL] = frac[:, 1:] - frac[:, :-1]
        inf[top_only] = 0.0
        inf[top_only, 0] = 1.0
        lay[:, L.INFRAC] = inf

    if beta is not None:
        rd = base.parametrizedSPAC.pars["root_distribution"]
        zmax = rd.get("z_rootMax_m", None) if isinstance(rd, dict) else None
        dz = base.parametrizedSPAC.soil_discretization["dz"]
        par = base.parametrizedSPAC.pars["params"]
        ub, inv = np.unique(np.asarray(beta, float), return_inverse=True)
        for j, b in enumerate(ub):
            prof, t_ini = root_growth_tini(THICK[0], Rootden_beta_(float(b), dz_m=dz, z_rootMax_m=zmax),
                                           par["INITRDEP"], par["RGRORATE"], par["RGROPER"])
            sel = inv == j
            lay[sel, L.RELDEN] = prof
            lay[sel, L.TINI] = t_ini
            finite = t_ini[t_ini < 1e19]
            const = par["RGROPER"] <= 0 or len(finite) == 0 or par["AGE_baseline_yrs"] > finite.max() + par["RGROPER"]
            scal[sel, S.RELDEN_CONST] = 1.0 if const else 0.0

    # initial state (init_LWFB90_u0!, src/func_DiffEq_definition_u0.jl:84-139)
    ix = lo.StateIndex(N, iso)
    u0 = np.repeat(base.u0[None], M, axis=0)
    if psim_init is not None or soil_changed:
        df = base.parametrizedSPAC.soil_discretization["df"]
        psim0 = (np.repeat(np.asarray(df["uAux_PSIM_init_kPa"], float)[None], M, axis=0) if psim_init is None
                 else np.repeat(np.asarray(psim_init, float)[:, None], N, axis=1))
        W0 = np.where(psim0 <= 0, (1 + (-ALPHA * psim0 / 9.81) ** NPAR) ** (-MPAR), 1.0)
        sw0 = FTheta_MvG(W0, THSAT, THR) * SWATMAX / THSAT
        u0[:, ix.SWATI_mm] = sw0
        new_swat = np.zeros(M)
        for i in range(N):                       # sequential sum like the reference's sum()
            new_swat = new_swat + sw0[:, i]
        u0[:, ix.acc("StorageSWAT")] = new_swat
        u0[:, ix.acc("StorageWATER")] = (u0[:, ix.INTR_mm] + u0[:, ix.INTS_mm] + u0[:, ix.SNOW_mm] + new_swat
                                         + u0[:, ix.GWAT_mm])
    return Ensemble(N=N, iso=iso, t0_day=base.t0_day, n_days=base.n_days, forcing=base.forcing[None].copy(),
                    site_of_member=np.zeros(M, np.int32), scal=scal, lay=lay, u0=u0)


def draw_knobs(base: DiscretizedSPAC, M: int, seed: int = 1234, first_member: int = 0, members=None) -> dict:
    """The perturbations of members first_member .. first_member+M-1 of the synthetic ensemble of SURVEY.md section
    8(d) - or of the members listed in `members` (global ids, any order: the generator is counter-based) - as keyword
    arguments of batched_setup (Philox4x32-10, key = seed, counter = global member index).  Global member 0 is the
    unmodified site definition (neutral knobs; `keep` marks it)."""
    idx = np.arange(first_member, first_member + M, dtype=np.uint64) if members is None else np.asarray(members, dtype=np.uint64)
    U = uniforms(idx, 20, seed)
    u = lambda j, a, b: a + (b - a) * U[:, j]
    logu = lambda j, a, b: np.exp(np.log(a) + (np.log(b) - np.log(a)) * U[:, j])
    S = lo.S
    b0 = base.scal
    keep = idx == 0

    def fix(arr, neutral):
        arr = np.asarray(arr, float).copy()
        arr[keep] = neutral
        return arr
    params = {
        "DRAIN": fix(u(5, 0.1, 1.0), b0[S.DRAIN]),
        "INFEXP": fix(u(6, 0.0, 1.0), float(base.parametrizedSPAC.pars["params"]["INFEXP"])),
        "GLMAX": fix(b0[S.GLMAX] * u(7, 0.7, 1.3), b0[S.GLMAX]),
        "CVPD": fix(b0[S.CVPD] * u(8, 0.7, 1.3), b0[S.CVPD]),
        "R5": fix(b0[S.R5] * u(9, 0.7, 1.3), b0[S.R5]),
        "MXKPL": fix(b0[S.MXKPL] * u(10, 0.7, 1.3), b0[S.MXKPL]),
        "PSICR": fix(b0[S.PSICR] * u(11, 0.7, 1.3), b0[S.PSICR]),
        "RSSA": fix(b0[S.RSSA] * u(12, 0.5, 2.0), b0[S.RSSA]),
        "ALB": fix(u(13, 0.1, 0.3), b0[S.ALB]),
        "CINTRL": fix(b0[S.CINTRL] * u(14, 0.7, 1.3), b0[S.CINTRL]),
        "MAXLAI": fix(b0[S.MAXLAI] * u(15, 0.7, 1.3), b0[S.MAXLAI]),
        "DISPERSIVITY": fix(logu(16, 5.0, 80.0) / 1000, b0[S.DISPERSIVITY]),
        "VXYLEM": fix(u(17, 5.0, 40.0), b0[S.VXYLEM]),
    }
    rd = base.parametrizedSPAC.pars["root_distribution"]
    beta = None
    if isinstance(rd, dict):
        # quantised to 64 levels so that the per-beta root profiles are shared between members
        levels = 0.85 + (0.99 - 0.85) * (np.floor(U[:, 18] * 64) + 0.5) / 64
        beta = fix(levels, rd["beta"])
    df = base.parametrizedSPAC.soil_discretization["df"]
    psim_uniform = np.all(df["uAux_PSIM_init_kPa"] == df["uAux_PSIM_init_kPa"][0])
    psim = fix(u(19, -12.0, -3.0), df["uAux_PSIM_init_kPa"][0]) if psim_uniform else None
    return dict(ths_f=fix(u(0, 0.85, 1.15), 1.0), ksat_f=fix(logu(1, 1 / 3, 3.0), 1.0),
                alpha_f=fix(logu(2, 0.5, 2.0), 1.0), nm1_f=fix(u(3, 0.8, 1.25), 1.0),
                gravel_f=fix(u(4, 0.8, 1.2), 1.0), params=params, beta=beta, psim_init=psim, keep=keep)


def make_ensemble(base: DiscretizedSPAC, M: int, seed: int = 1234, first_member: int = 0) -> Ensemble:
    """Members first_member .. first_member+M-1 of the synthetic ensemble of SURVEY.md section 8(d), built on the host
    (numpy).  Global member 0 is the unmodified site definition."""
    kn = draw_knobs(base, M, seed, first_member)
    keep = kn.pop("keep")
    ens = batched_setup(base, **kn)
    if keep.any():   # member 0 must be bit-identical to the base definition
        j = int(np.nonzero(keep)[0][0])
        ens.scal[j], ens.lay[j], ens.u0[j] = base.scal, base.lay, base.u0
    return ens


@dataclass
class KnobBlock:
    """What lwfb_setup_members (include/lwfb90_setup.h) takes instead of finished member blocks: the base member and
    one row of knob values per member - the batched setup then runs on the GPU."""
    N: int
    iso: bool
    t0_day: int
    n_days: int
    forcing: np.ndarray          # [n_sites, n_rows, F]
    base_scal: np.ndarray        # [PS]
    base_lay: np.ndarray         # [PL, N]
    base_u0: np.ndarray          # [L]
    knobs: np.ndarray            # [M, K + n_over], NaN = keep the base value
    over_idx: np.ndarray         # [n_over] int32, LWFB_S_* indices
    profiles: np.ndarray         # [n_profiles, 2N+1]
    psim_base: np.ndarray        # [N]
    ilayer: int

    @property
    def M(self) -> int:
        return self.knobs.shape[0]


K_THS_F, K_KSAT_F, K_ALPHA_F, K_NM1_F, K_GRAVEL_F, K_INFEXP, K_ROOTPROFILE, K_PSIM_INIT, K_FIXED = range(9)


def knob_block(base: DiscretizedSPAC, *, ths_f=None, ksat_f=None, alpha_f=None, nm1_f=None, gravel_f=None,
               params=None, beta=None, psim_init=None, keep=None) -> KnobBlock:
    """Same keyword arguments as batched_setup; returns the inputs of the DEVICE-side setup.  Members flagged in `keep`
    get an all-NaN row (= the base member, bit for bit)."""
    S = lo.S
    some = [x for x in (ths_f, ksat_f, alpha_f, nm1_f, gravel_f, beta, psim_init) if x is not None]
    params = dict(params or {})
    some += list(params.values())
    M = len(some[0]) if some else 1
    infexp = params.pop("INFEXP", None)
    names = list(params)
    kb = np.full((M, K_FIXED + len(names)), np.nan)
    for col, v in ((K_THS_F, ths_f), (K_KSAT_F, ksat_f), (K_ALPHA_F, alpha_f), (K_NM1_F, nm1_f), (K_GRAVEL_F, gravel_f),
                   (K_INFEXP, infexp), (K_PSIM_INIT, psim_init)):
        if v is not None:
            kb[:, col] = np.asarray(v, float)
    for j, name in enumerate(names):
        kb[:, K_FIXED + j] = np.asarray(params[name], float)
    N = base.N
    profiles = np.zeros((0, 2 * N + 1))
    if beta is not None:
        rd = base.parametrizedSPAC.pars["root_distribution"]
        zmax = rd.get("z_rootMax_m", None) if isinstance(rd, dict) else None
        dz = base.parametrizedSPAC.soil_discretization["dz"]
        par = base.parametrizedSPAC.pars["params"]
        THICK = base.lay[lo.L.THICK]
        ub, inv = np.unique(np.asarray(beta, float), return_inverse=True)
        profiles = np.zeros((len(ub), 2 * N + 1))
        for j, b in enumerate(ub):
            prof, t_ini = root_growth_tini(THICK, Rootden_beta_(float(b), dz_m=dz, z_rootMax_m=zmax),
                                           par["INITRDEP"], par["RGRORATE"], par["RGROPER"])
            finite = t_ini[t_ini < 1e19]
            const = par["RGROPER"] <= 0 or len(finite) == 0 or par["AGE_baseline_yrs"] > finite.max() + par["RGROPER"]
            profiles[j, :N], profiles[j, N:2 * N], profiles[j, 2 * N] = prof, t_ini, 1.0 if const else 0.0
        kb[:, K_ROOTPROFILE] = inv
    if keep is not None:
        kb[np.asarray(keep, bool)] = np.nan
    df = base.parametrizedSPAC.soil_discretization["df"]
    return KnobBlock(N=N, iso=base.iso, t0_day=base.t0_day, n_days=base.n_days, forcing=base.forcing[None].copy(),
                     base_scal=base.scal.copy(), base_lay=base.lay.copy(), base_u0=base.u0.copy(), knobs=kb,
                     over_idx=np.array([getattr(S, n) for n in names], dtype=np.int32), profiles=profiles,
                     psim_base=np.asarray(df["uAux_PSIM_init_kPa"], float).copy(), ilayer=int(base.ILAYER))


def make_knob_block(base: DiscretizedSPAC, M: int, seed: int = 1234, first_member: int = 0, members=None) -> KnobBlock:
    """The synthetic ensemble of make_ensemble as a knob block for the device-side setup (members: explicit global ids)."""
    return knob_block(base, **draw_knobs(base, M, seed, first_member, members))
