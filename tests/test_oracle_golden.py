"""The CPU oracle (mode A, host libm) against every golden vector the reference holds for this path
(G1..G6).  These tests pin the oracle; the GPU tests then pin the CUDA path to the oracle."""
import json
import os

import numpy as np
import pytest

from conftest import GOLD, inputs, outputs, read_table

import oracle as orc
from lwfbrook90.jl_b200 import api, layout as lo


def run(sim, n_days=None, mode="libm", **kw):
    orc.use(mode)
    out, st, cnt, nt = orc.run_member(sim.N, sim.iso, sim.t0_day, sim.n_days if n_days is None else n_days,
                                      sim.forcing, sim.scal, sim.lay, sim.u0, **kw)
    assert st == 0
    return out, cnt, nt


# ---------------------------------------------------------------- G6: leaf known answers
def test_kkmean_known_answers():            # test/01-unit-tests.jl:41-47
    L = orc.lib()
    assert np.isclose(L.lwfb_oracle_kkmean(50., 50., 1., 1.), 50)
    assert np.isclose(L.lwfb_oracle_kkmean(50., 50., 1., 5.), 50)
    assert np.isclose(L.lwfb_oracle_kkmean(10., 100., 1., 1.), 31.6227766)
    assert np.isclose(L.lwfb_oracle_kkmean(10., 100., 1., 5.), 14.6779926)
    assert np.isclose(L.lwfb_oracle_kkmean(10., 100., 5., 1.), 68.1292069)


def test_esat_pm_known_answers():           # test/01-unit-tests.jl:50-66
    import ctypes as C
    L = orc.lib()
    es, de = C.c_double(), C.c_double()
    L.lwfb_oracle_esat(0.0, C.byref(es), C.byref(de))
    assert np.isclose(es.value, 0.61078) and np.isclose(de.value, 0.044448927)
    L.lwfb_oracle_esat(25.0, C.byref(es), C.byref(de))
    assert np.isclose(es.value, 3.1674898302368564) and np.isclose(de.value, 0.18866467946037985)
    assert np.isclose(L.lwfb_oracle_pm(150., -0.43, 0.06, 146., 19640.), 0.58512224)
    assert np.isclose(L.lwfb_oracle_pm(150., 0.30, 0.06, 23.8, 578.), 14.0411894)


def test_fastpow_is_a_rough_power():
    # DiffEqBase.fastpow is a ~1e-4 accurate Float32 approximation; the controller depends on it
    L = orc.lib()
    for x in (1e-4, 0.03, 0.5, 1.0, 7.0):
        for y in (0.14, 0.08):
            assert abs(L.lwfb_oracle_fastpow(x, y) / x ** y - 1) < 2e-3


def test_tsit5_tableau_consistency():
    # row sums equal the nodes, weights sum to one, error weights sum to zero
    a = {2: [0.161], 3: [-0.008480655492356989, 0.335480655492357],
         4: [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
         5: [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
         6: [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383],
         7: [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774]}
    c = {2: 0.161, 3: 0.327, 4: 0.9, 5: 0.9800255409045097, 6: 1.0, 7: 1.0}
    for s, row in a.items():
        assert abs(sum(row) - c[s]) < 1e-14
    bt = [-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
          0.5823571654525552, -0.45808210592918697, 0.015151515151515152]
    assert abs(sum(bt)) < 1e-15
    b = a[7]
    cs = [0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0]
    for p in range(1, 5):                    # quadrature conditions up to order 5
        assert abs(sum(bi * ci ** p for bi, ci in zip(b, cs)) - 1 / (p + 1)) < 1e-13


# ---------------------------------------------------------------- G1: DAV2020-full run_example
def test_G1_dav2020_full_states(lw):
    sim = lw.setup(lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True), requested_tspan=(0, 300))
    out, cnt, _ = run(sim)
    ix = lo.StateIndex(sim.N, True)
    h, a = read_table(outputs("DAV2020-full_u_sol_referenceu_belowground.csv"))
    days = a[:, 0].astype(int)
    sw = np.array([out[k, ix.SWATI_mm] for k in days])
    assert np.max(np.abs(sw - a[:, 1:6]) / a[:, 1:6]) < 1e-12
    assert np.max(np.abs(np.array([out[k, ix.SWATI_d18O] for k in days]) - a[:, 6:11])) < 1e-7
    assert np.max(np.abs(np.array([out[k, ix.SWATI_d2H] for k in days]) - a[:, 11:16])) < 1e-6
    th_psi = [sim.p_soil.soilvar(out[k, ix.SWATI_mm]) for k in days]
    assert np.max(np.abs(np.array([x[3] for x in th_psi]) - a[:, 16:21]) / a[:, 16:21]) < 1e-12
    assert np.max(np.abs(np.array([x[1] for x in th_psi]) - a[:, 21:26]) / np.abs(a[:, 21:26])) < 1e-11
    h, a = read_table(outputs("DAV2020-full_u_sol_referenceu_mm.csv"))
    mine = np.array([[out[int(k), getattr(ix, n)] for n in ("GWAT_mm", "INTS_mm", "INTR_mm", "SNOW_mm", "CC", "SNOWLQ", "RWU_mm", "XYLEM_mm")]
                     for k in a[:, 0]])
    assert np.max(np.abs(mine - a[:, 1:])) < 1e-12
    h, a = read_table(outputs("DAV2020-full_u_sol_referenceu_δ.csv"))
    cols = ["GWAT", "INTS", "INTR", "SNOW", "RWU", "XYLEM"]
    mine = np.array([[out[int(k), getattr(ix, n + "_d18O")] for n in cols] + [out[int(k), getattr(ix, n + "_d2H")] for n in cols]
                     for k in a[:, 0]])
    ref = np.concatenate([a[:, 2:8], a[:, 9:15]], axis=1)
    assert np.array_equal(np.isnan(mine), np.isnan(ref))        # NaN semantics are part of the contract
    assert np.nanmax(np.abs(mine - ref)) < 1e-6
    # the reference's developer comment for this run class: steps sit at the dtmax floor
    assert 24 <= cnt[0] / 300 < 26


# ---------------------------------------------------------------- G2/G3: INFEXP1
def _flux_cmp(out, ix, path):
    h, a = read_table(path)
    h = h[1:]
    a = a[:, 1:]
    mine = np.array([[out[k, ix.acc(n)] for n in h] for k in range(a.shape[0])])
    return h, mine, a


@pytest.mark.parametrize("modified", [False, True])
def test_G2_G3_infexp1(lw, infexp1, modified):
    sim = infexp1
    if modified:
        sim = lw.remakeSPAC(sim, params=dict(BYPAR=1, QDEPTH_m=0.40, QFFC=0.2, QFPAR=0.3, GSC=0.05, GSP=0.2))
    assert sim.N == 42
    out, cnt, _ = run(sim)
    name = "INFEXP1-modified_fluxes_referencedf.csv" if modified else "INFEXP1_fluxes_referencedf.csv"
    h, mine, ref = _flux_cmp(out, lo.StateIndex(42, False), outputs(name))
    # before the profile saturates (day 231) the run is well conditioned: round-off level agreement
    assert np.max(np.abs(mine[:232] - ref[:232])) < 1e-7
    # afterwards the reference algorithm is chaotic (saturation clamps + flux limiter, the reference's own CI
    # skips 12 of these series); storages still agree to a fraction of a per cent
    for c in ("StorageSWAT", "StorageWATER"):
        j = h.index(c)
        assert np.max(np.abs(mine[:, j] - ref[:, j]) / ref[:, j]) < 2e-3
    for c in ("cum_d_prec", "cum_d_rfal", "cum_d_sfal", "cum_d_rint", "cum_d_pint", "cum_d_irvp"):
        j = h.index(c)
        assert np.allclose(mine[:, j], ref[:, j], rtol=1e-4, atol=1e-4)


def test_infexp1_is_chaotic_after_saturation(infexp1):
    """Documents why: a 1e-15 relative change of KSAT changes the bypass flow of day 232 in the 2nd digit."""
    sim = infexp1
    ix = lo.StateIndex(42, False)
    a, _, _ = run(sim, n_days=240)
    lay = sim.lay.copy()
    lay[lo.L.KSAT] *= (1 + 1e-15)
    orc.use("libm")
    b = orc.run_member(sim.N, False, 0, 240, sim.forcing, sim.scal, lay, sim.u0)[0]
    assert np.max(np.abs(a[:231] - b[:231])) < 1e-9
    assert abs(a[232, ix.acc("byfl")] - b[232, ix.acc("byfl")]) > 1e-3


# ---------------------------------------------------------------- G4: DAV2020-full modified
def _dav_modified(lw):
    m3 = lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True,
                     canopy_evolution=dict(DENSEF_rel=100, HEIGHT_rel=100, SAI_rel=100,
                                           LAI_rel=dict(DOY_Bstart=120, Bduration=21, DOY_Cstart=270, Cduration=60,
                                                        LAI_perc_BtoC=100, LAI_perc_CtoB=20)),
                     Δz_thickness_m=[0.02] * 40, root_distribution=dict(beta=0.98, z_rootMax_m=-0.5),
                     IC_soil=dict(PSIM_init_kPa=-7.0, delta18O_init_permil=-9.0, delta2H_init_permil=-11.0),
                     IC_scalar=dict(amount=dict(u_GWAT_init_mm=0, u_INTS_init_mm=0, u_INTR_init_mm=0, u_SNOW_init_mm=0,
                                                u_CC_init_MJ_per_m2=0, u_SNOWLQ_init_mm=0),
                                    d18O=dict(u_GWAT_init_permil=-13., u_INTS_init_permil=-13., u_INTR_init_permil=-13., u_SNOW_init_permil=-13.),
                                    d2H=dict(u_GWAT_init_permil=-95., u_INTS_init_permil=-95., u_INTR_init_permil=-95., u_SNOW_init_permil=-95.)),
                     storm_durations_h=[4, 4, 4, 4, 8, 8, 8, 4, 4, 4, 4, 4])
    s3a = lw.setup(m3, requested_tspan=(0, 300))
    return lw.remakeSPAC(s3a, soil_horizons=dict(ths_=0.4, Ksat_mmday=3854.9, alpha_per_m=7.11, gravel_volFrac=0.1),
                         LAI_rel=dict(DOY_Bstart=115), root_distribution=dict(beta=0.88, z_rootMax_m=-0.6),
                         params=dict(IDEPTH_m=0.50, INFEXP=0.33, BYPAR=1, QDEPTH_m=0.70, QFFC=0.2, QFPAR=0.3, IMPERV=0.01,
                                     DSLOPE=10, DRAIN=.33, GSC=0.05, GSP=0.2, ALB=0.15, ALBSN=0.7, RSSA=720., PSICR=-1.6,
                                     FXYLEM=0.4, MXKPL=16.5, MAXLAI=9.999, GLMAX=.00801, R5=235., CVPD=1.9, CINTRL=0.18))


def test_G4_dav2020_modified_fluxes(lw):
    sim = _dav_modified(lw)
    assert sim.N == 40 and sim.tspan == (0.0, 300.0)
    out, cnt, _ = run(sim)
    h, mine, ref = _flux_cmp(out, lo.StateIndex(40, True), outputs("DAV2020-full-modified_fluxes_referencedf.csv"))
    # the reference's own tolerance (test/03-regression-tests.jl:318-348), applied to ALL 30 series
    assert np.all(np.abs(mine - ref) <= 1e-4 + 1e-4 * np.abs(ref))
    assert np.max(np.abs(mine - ref)) < 1e-5


# ---------------------------------------------------------------- G5: inline tables of the unit tests
def _g5_model(lw):
    """Run definition of the unit tests' inline tables, test/01-unit-tests.jl:858-875."""
    m5 = lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True, Δz_thickness_m=[0.05] * 22,
                     root_distribution=dict(beta=0.77, z_rootMax_m=-0.5),
                     IC_soil=dict(PSIM_init_kPa=-7.0, delta18O_init_permil=-10.11111, delta2H_init_permil=-91.1111),
                     canopy_evolution=dict(DENSEF_rel=100, HEIGHT_rel=100, SAI_rel=100,
                                           LAI_rel=dict(DOY_Bstart=120, Bduration=21, DOY_Cstart=270, Cduration=60,
                                                        LAI_perc_BtoC=95, LAI_perc_CtoB=70)))
    sim = lw.setup(m5)
    assert sim.N == 22 and sim.tspan == (0.0, 365.0)
    return sim


@pytest.fixture(scope="module")
def g5(lw):
    sim = _g5_model(lw)
    cold, cnt, nt = run(sim)
    # the golden comes from the SECOND simulate! on the same problem: warm du_NTFLI cache
    warm, cnt2, _ = run(sim, ntfli_init=nt)
    sim.solution = api.ODESolution(sim, np.arange(366.0), warm, 0, cnt2)
    import datetime as dt
    sim.ODESolution_datetime = [sim.parametrizedSPAC.reference_date + dt.timedelta(days=k) for k in range(366)]
    return sim, json.load(open(os.path.join(GOLD, "golden_inline.json")))


def test_G5_initial_theta(g5):              # test/01-unit-tests.jl:887-893
    sim, _ = g5
    t = api.get_soil_("θ", sim, depths_to_read_out_mm=[100, 1000, 200, 300, 400, 150], days_to_read_out_d=0)
    for k, v in t.items():
        if k != "time":
            assert np.isclose(v[0], 0.20058687988)
    assert api.get_soil_idx(sim, [100, 1000, 1200, 200, 300, 400, 150]) == {150: 3, 100: 2, 200: 4, 300: 6, 400: 8, 1000: 21, 1200: 0}


def test_G5_isotope_table(g5):              # test/01-unit-tests.jl:951-1014
    sim, gold = g5
    ref = np.array(gold["reference_isotopePlotPermutedDims"])
    sol, ix = sim.solution, sim.solution.ix
    days = [0, 9, 99, 299]
    rec = np.array([sol.record_at(d) for d in days])
    assert np.allclose(ref[0], days)
    names = ["INTS", "INTR", "SNOW", "GWAT", "RWU", "XYLEM"]
    for j, n in enumerate(names):
        assert np.allclose(rec[:, getattr(ix, n + "_d18O")], ref[6 + j], rtol=0, atol=1e-6)
        assert np.allclose(rec[:, getattr(ix, n + "_d2H")], ref[12 + j], rtol=0, atol=1e-5)
    assert np.max(np.abs(rec[:, ix.SWATI_d18O].T - ref[19:41])) < 1e-6
    assert np.max(np.abs(rec[:, ix.SWATI_d2H].T - ref[41:63])) < 1e-5
    # precipitation delta at the read-out days
    assert np.allclose([sol.p_d18O_PREC(d) for d in days], ref[4]) and np.allclose([sol.p_d2H_PREC(d) for d in days], ref[5])


def test_G5_water_partitioning(g5):         # test/01-unit-tests.jl:907-945
    """The reference compares whole matrices with isapprox(atol=1e-5, rtol=1e-5), i.e. in the Frobenius
    norm: ||A-B|| <= max(atol, rtol*max(||A||,||B||)).  Same criterion here, plus an element-wise bound.
    (The golden literals carry 6 significant digits; the largest element-wise deviation, 2.7e-4 relative,
    is the transpiration deficit Td of the one supply-limited day of the year, 2021-06-29.)"""
    sim, gold = g5
    daily, monthly, yearly = api.get_water_partitioning(sim)
    cols = ["ETa", "Esoil", "Esnow", "Einterception", "Ta", "Precip", "Td", "D", "R", "Swat"]

    def check(mine, ref):
        assert mine.shape == ref.shape
        assert np.linalg.norm(mine - ref) <= max(1e-5, 1e-5 * max(np.linalg.norm(mine), np.linalg.norm(ref)))
        big = np.abs(ref) > 1e-3
        assert np.max(np.abs(mine - ref)[big] / np.abs(ref)[big]) < 3e-4
        assert np.max(np.abs(mine - ref)[~big]) < 1e-6
    rows = [9, 19, 99, 179, 249, 259]       # Julia rows 10, 20, 100, 180, 250, 260
    check(np.array([[daily[c][r] for c in cols] for r in rows]), np.array(gold["reference_daily_check"]))
    check(np.array([[monthly[c][r] for c in cols] for r in range(len(monthly["year"]))]), np.array(gold["reference_monthly_check"]))
    check(np.array([[yearly[c][r] for c in cols] for r in range(len(yearly["year"]))]), np.array(gold["reference_yearly_check"]))


# ---------------------------------------------------------------- oracle mode B vs mode A
def test_elem_functions_accuracy():
    """include/lwfb90_elem.h against the host libm (<= 3 ulp; pow ~ |y log x| * 1.2e-16)."""
    import math
    rng = np.random.default_rng(3)
    orc.use("elem")
    cases = {0: (rng.uniform(-700, 700, 4000), math.exp), 1: (10 ** rng.uniform(-300, 300, 4000), math.log),
             3: (rng.uniform(-50, 50, 4000), math.sin), 4: (rng.uniform(-50, 50, 4000), math.cos),
             6: (rng.uniform(-1, 1, 4000), math.asin), 7: (rng.uniform(-1, 1, 4000), math.acos)}
    for fn, (x, f) in cases.items():
        got = orc.elem(fn, x)
        ref = np.array([f(v) for v in x])
        ulp = np.abs(got - ref) / np.spacing(np.abs(ref))
        assert ulp.max() <= 3, (fn, ulp.max())
    x = 10 ** rng.uniform(-6, 0, 4000)
    y = rng.uniform(-25, 25, 4000)
    got = orc.elem(2, x, y)
    assert np.max(np.abs(got / x ** y - 1)) < 1e-13
    assert orc.elem(2, [0.0, 0.0, 3.0], [0.5, 0.0, 0.0]).tolist() == [0.0, 1.0, 1.0]
    orc.use("libm")


def test_mode_B_tracks_mode_A_on_a_well_conditioned_run(dav_full):
    a, _, _ = run(dav_full, n_days=120, mode="libm")
    b, _, _ = run(dav_full, n_days=120, mode="elem")
    ix = lo.StateIndex(dav_full.N, True)
    assert np.max(np.abs(a[:, ix.SWATI_mm] - b[:, ix.SWATI_mm]) / a[:, ix.SWATI_mm]) < 1e-10
    assert np.nanmax(np.abs(a[:, ix.SWATI_d18O] - b[:, ix.SWATI_d18O])) < 1e-6
    orc.use("libm")
