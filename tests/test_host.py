"""Host-side twin of loadSPAC / setup / remakeSPAC / ensemble generation against the reference's
unit-test known answers (test/01-unit-tests.jl)."""
import numpy as np
import pytest

from conftest import inputs, outputs

from lwfbrook90.jl_b200 import ensemble as en, layout as lo
from lwfbrook90.jl_b200.discretize import Rootden_beta_
from lwfbrook90.jl_b200.kpt import SoilPar


def test_layout_matches_header():
    assert lo.NF == 16 and lo.NACCUM == 31 and lo.NPL == 18
    assert lo.F.GLOBRAD == 0 and lo.F.D2H == 13
    assert lo.state_len(5, True) == 96 and lo.state_len(7, False) == 74 and lo.state_len(42, False) == 249
    ix = lo.StateIndex(5, True)
    assert ix.GWAT_mm == 15 and ix.RWU_mm == 27 and ix.CC == 33 and ix.accum0 == 65 and ix.L == 96
    assert lo.ACCUM_NAMES[lo.A.FLOW] == "flow" and lo.ACCUM_NAMES[lo.A.BALERD_TOTAL] == "BALERD_total"


def test_kpt_soilpar_known_answers():        # test/01-unit-tests.jl:68-112
    kw = dict(STONEF=[0.010, 0.175, 0.175], THSAT=[0.714, 0.668, 0.656], KSAT=[24864., 12881., 10516.],
              alpha=[1147., 1274., 1215.], n=[1.051225, 1.051052, 1.051055],
              m=[0.048728863944445866, 0.04857228757473475, 0.04857500321105945], l=[4.6703, 4.4782, 4.5016],
              thr=[0.069, 0.069, 0.069], Kfc=2.0)
    p1 = SoilPar(THICK=[40., 40., 120.], **kw)
    assert np.allclose(p1.p_PSIF, [-0.03210102700011719, -0.02096528992202748, -0.019805546964157438], rtol=1.5e-8)
    assert np.allclose(p1.p_THETAF, [0.6652527196951767, 0.6299247343092094, 0.6208286442505493], rtol=1.5e-8)
    assert np.allclose(p1.p_PSIG, [-0.19619999999999999, -0.5886, -1.3734], rtol=1.5e-8)
    assert np.allclose(p1.p_SWATMAX, [28.2744, 22.044, 64.944], rtol=1.5e-8)
    assert np.allclose(p1.p_WETF, [0.924422821232832, 0.9364352826531043, 0.9400828692513616], rtol=1.5e-8)
    p2 = SoilPar(THICK=[120., 120., 360.], **kw)
    assert np.allclose(p2.p_PSIG, [-0.5886, -1.7658, -4.1202000000000005], rtol=1.5e-8)
    assert np.allclose(p2.p_SWATMAX, p1.p_SWATMAX * 3)


@pytest.mark.parametrize("zmax", [None, -0.5])
def test_rootden_beta_known_answers(lw, zmax):       # test/01-unit-tests.jl:384-421
    expect = {
        (4, None): (slice(None), [0.7201999283091463, 0.25792436725268547, 0.020201681877684945, 0.0015544179126264956, 0.00011960464785673769]),
        (100, None): (slice(0, 100, 10), [0.09207591746299604, 0.03520806629944542, 0.01262159612357775, 0.004524664528628884,
                                          0.0016220285371347674, 0.0005814743963076271, 0.00020845038531732468,
                                          7.472652865688285e-5, 2.6788408553974765e-5, 9.603267350270948e-6]),
        (200, None): (slice(0, 200, 10), [0.04823125873564055, 0.030012878588235002, 0.017969819063652963, 0.010759194464839009,
                                          0.006441927162548269, 0.0038570197521006896, 0.002309340200954726, 0.0013826872835797437,
                                          0.0008278659520944154, 0.0004956739262566578, 0.0002967782894672004, 0.00017769212466797057,
                                          0.00010639083885043477, 6.370012521630651e-5, 3.8139618001095856e-5, 2.2835598145652136e-5,
                                          1.367251613943244e-5, 8.186240464953907e-6, 4.901404559846817e-6, 2.93465196418323e-6]),
        (4, -0.5): (slice(None), [0.73630716625382, 0.26369283374617997, 0.0, 0.0, 0.0]),
        (100, -0.5): (slice(0, 50, 5), [0.09899773754146445, 0.06322449003427368, 0.0378548376480284, 0.02266508962874636,
                                        0.013570426391879014, 0.0081251155620333, 0.0, 0.0, 0.0, 0.0]),
        (200, -0.5): (slice(0, 100, 10), [0.05204344024903263, 0.032385085823067206, 0.01939014716267554, 0.011609597363562048,
                                          0.006951094791249151, 0.0041618772196682025, 0.0, 0.0, 0.0, 0.0]),
    }
    for dz in ([0.5] * 4, [0.02] * 100, [0.01] * 200):
        rd = dict(beta=0.95) if zmax is None else dict(beta=0.95, z_rootMax_m=zmax)
        m = lw.loadSPAC(inputs("Hammel-2001"), "Hammel_loam-NLayer-27-RESET=FALSE", simulate_isotopes=False,
                        Δz_thickness_m=dz, root_distribution=rd,
                        IC_soil=dict(PSIM_init_kPa=-7.0, delta18O_init_permil=-9.0, delta2H_init_permil=-11.0))
        s = lw.setup(m, ε=0.005)
        sl, ref = expect[(len(dz), zmax)]
        got = s.parametrizedSPAC.soil_discretization["df"]["Rootden_"][sl]
        assert np.allclose(got, ref, rtol=1.5e-8, atol=0)
    assert np.isclose(Rootden_beta_(0.95, Δz_m=np.full(100, 0.02)).sum(), 1.0)


def test_dav_full_setup_matches_golden_row0(dav_full):
    s = dav_full
    assert s.N == 5 and s.iso and s.tspan == (0.0, 365.0) and s.forcing.shape == (366, 16)
    ix = lo.StateIndex(5, True)
    # examples/DAV2020-full_u_sol_referenceu_belowground.csv row t = 0
    assert np.allclose(s.u0[ix.SWATI_mm], [5.08407918140519, 8.987008653999073, 38.51575137428175, 20.541734066283595,
                                           15.406300549712698], rtol=1e-14)
    assert s.u0[ix.XYLEM_mm] == 5.0 and s.u0[ix.GWAT_mm] == 1.0 and s.u0[ix.INTS_d18O] == -13.0
    assert np.all(s.u0[ix.aux_theta] == 0) and s.ILAYER == 3 and s.QLAYER == 0
    assert np.isclose(s.lay[lo.L.INFRAC].sum(), 1.0) and np.all(s.lay[lo.L.INFRAC][3:] == 0)
    # delta of precipitation: first sample overwritten by the second, knots at sample day + 0.5
    assert s.forcing[0, lo.F.D18O] == -15.04 and s.forcing[12, lo.F.D18O] == -15.04 and s.forcing[13, lo.F.D18O] == -20.19
    assert s.forcing[0, lo.F.DOY] == 1 and s.forcing[365, lo.F.DOY] == 1 and s.forcing[0, lo.F.IHD] == 2


def test_bare_minimum_refinement(lw):        # test/01-unit-tests.jl:423-470
    m = lw.loadSPAC(inputs("DAV2020-bare-minimum"), "DAV2020-minimal", simulate_isotopes=True, Δz_thickness_m=[0.1] * 11,
                    root_distribution=dict(beta=0.77, z_rootMax_m=-0.5),
                    IC_soil=dict(PSIM_init_kPa=-7.0, delta18O_init_permil=-10.11111, delta2H_init_permil=-91.1111),
                    canopy_evolution=dict(DENSEF_rel=100, HEIGHT_rel=100, SAI_rel=100,
                                          LAI_rel=dict(DOY_Bstart=120, Bduration=21, DOY_Cstart=270, Cduration=60,
                                                       LAI_perc_BtoC=95, LAI_perc_CtoB=70)),
                    storm_durations_h=[5.44] * 12,
                    IC_scalar=dict(amount=dict(u_GWAT_init_mm=1., u_INTS_init_mm=13.7, u_INTR_init_mm=0., u_SNOW_init_mm=22.222,
                                               u_CC_init_MJ_per_m2=0.101010, u_SNOWLQ_init_mm=0.),
                                   d18O=dict(u_GWAT_init_permil=-11.111, u_INTS_init_permil=-12.222, u_INTR_init_permil=-13.333, u_SNOW_init_permil=-14.444),
                                   d2H=dict(u_GWAT_init_permil=-95.111, u_INTS_init_permil=-95.222, u_INTR_init_permil=-95.333, u_SNOW_init_permil=-95.444)))
    with pytest.raises(ValueError, match="goes beyond input forcing data"):
        lw.setup(m, requested_tspan=(0., 400))
    with pytest.raises(AssertionError):
        lw.setup(m, soil_output_depths_m=[-1.0755, -1.096])
    s = lw.setup(m, soil_output_depths_m=[-1.0755, -1.096], ε=0.005)
    itf = np.cumsum(s.p_soil.p_THICK)
    assert np.isclose(m.pars["params"]["IDEPTH_m"], 0.333)
    assert np.isclose(s.p_soil.p_THICK[s.lay[lo.L.INFRAC] != 0].sum(), 333.0)
    assert np.any(np.abs(itf - 333.0) < 1e-5) and np.any(np.abs(itf - 30.0) < 1e-5)
    assert np.any(np.abs(itf - 1075.5) < 1e-5) and np.any(np.abs(itf - 1080.5) < 1e-5) and np.any(np.abs(itf - 1096.0) < 1e-5)
    s0 = lw.setup(m)
    assert s0.N == 13                       # 11 layers + interfaces at 0.03 m and 0.333 m
    ix = lo.StateIndex(s0.N, True)
    assert s0.u0[ix.INTS_mm] == 13.7 and s0.u0[ix.SNOW_d2H] == -95.444 and s0.u0[ix.CC] == 0.101010


def test_remake_propagates(lw, dav_full):    # test/01-unit-tests.jl:533-759 (spirit)
    s = dav_full
    r = lw.remakeSPAC(s, params=dict(DRAIN=.33, BYPAR=1, IDEPTH_m=0.67, INFEXP=0.33, ALB=0.15),
                      soil_horizons=dict(ths_=0.4, Ksat_mmday=3854.9))
    assert r.scal[lo.S.DRAIN] == .33 and r.scal[lo.S.BYPAR] == 1 and r.scal[lo.S.ALB] == 0.15
    assert np.isclose(r.lay[lo.L.THSAT][0], 0.4) and np.isclose(r.lay[lo.L.KSAT][0], 3854.9)
    assert r.N == s.N + 1                     # IDEPTH_m = 0.67 needs a new interface
    assert np.allclose(r.lay[lo.L.THSAT], 0.4)  # all horizons scaled proportionally to the top one
    assert np.allclose(r.lay[lo.L.STONEF][[0, -1]], s.lay[lo.L.STONEF][[0, -1]])
    assert np.isclose(np.cumsum(r.p_soil.p_THICK)[r.lay[lo.L.INFRAC] != 0][-1], 670.0)
    assert s.scal[lo.S.DRAIN] == 0.51         # the original is untouched
    with pytest.raises(AssertionError):
        lw.remakeSPAC(s, params=dict(NOT_A_PARAM=1))
    with pytest.raises(AssertionError):
        lw.remakeSPAC(s, LAI_rel=dict(DOY_Bstart=115))   # canopy of this SPAC comes from meteoveg.csv


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    z = en.philox4x32_10(np.zeros((1, 4), np.uint32), (0, 0))[0]
    assert [hex(v) for v in z] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = en.philox4x32_10(np.full((1, 4), 0xFFFFFFFF, np.uint32), (0xFFFFFFFF, 0xFFFFFFFF))[0]
    assert [hex(v) for v in f] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_ensemble_member0_and_batched_setup(lw, dav_full):
    s = dav_full
    e = en.make_ensemble(s, 50)
    assert np.array_equal(e.scal[0], s.scal) and np.array_equal(e.lay[0], s.lay) and np.array_equal(e.u0[0], s.u0)
    # any block of members can be regenerated independently (counter-based RNG)
    e2 = en.make_ensemble(s, 10, first_member=20)
    assert np.array_equal(e2.scal, e.scal[20:30]) and np.array_equal(e2.lay, e.lay[20:30]) and np.array_equal(e2.u0, e.u0[20:30])
    # the vectorised setup agrees with remakeSPAC -> setup member by member
    j = 7
    U = en.uniforms(np.array([j], dtype=np.uint64), 20)[0]
    hz = s.parametrizedSPAC.pars["soil_horizons"].shp[0]
    lu = lambda x, a, b: np.exp(np.log(a) + (np.log(b) - np.log(a)) * x)
    r = lw.remakeSPAC(s, soil_horizons=dict(ths_=hz.p_THSAT * (0.85 + 0.3 * U[0]), Ksat_mmday=hz.p_KSAT * lu(U[1], 1 / 3, 3.0),
                                            alpha_per_m=hz.p_MvGalpha * lu(U[2], 0.5, 2.0),
                                            npar_=1 + (hz.p_MvGn - 1) * (0.8 + 0.45 * U[3]),
                                            gravel_volFrac=hz.p_STONEF * (0.8 + 0.4 * U[4])),
                      params=dict(DRAIN=0.1 + 0.9 * U[5], INFEXP=U[6]))
    for nm in ("THSAT", "KSAT", "ALPHA", "NPAR", "MPAR", "STONEF", "SWATMAX", "WETF", "PSIF", "THETAF", "INFRAC"):
        a, b = e.lay[j, getattr(lo.L, nm)], r.lay[getattr(lo.L, nm)]
        assert np.allclose(a, b, rtol=1e-13, atol=0), nm
    assert np.isclose(e.scal[j, lo.S.DRAIN], r.scal[lo.S.DRAIN], rtol=1e-15)


def test_shard_range():
    from lwfbrook90.jl_b200.parallel import shard_range
    for M in (1, 7, 8, 1000003):
        for W in (1, 2, 4, 8):
            r = [shard_range(M, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == M and all(r[k][1] == r[k + 1][0] for k in range(W - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_golden_csv_writers_match_reference_layout(lw, tmp_path):
    """the tables of the reference's regression tests (test/03-regression-tests.jl:35-62, :231-273)
    written in the goldens' CSV layout.  The solution comes from the CPU oracle here (the same ODESolution object that
    simulate_ builds from the GPU record): headers must equal the reference's golden files column for column, values
    agree at the oracle's pinned tolerances, and a write -> read round trip is exact."""
    import oracle as orc
    from lwfbrook90.jl_b200 import api, golden_io as gio
    sim = lw.setup(lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True), requested_tspan=(0, 300))
    orc.use("libm")
    out, st, cnt, _ = orc.run_member(sim.N, sim.iso, sim.t0_day, sim.n_days, sim.forcing, sim.scal, sim.lay, sim.u0)
    sim.solution = api.ODESolution(sim, sim.t0_day + np.arange(sim.n_days + 1, dtype=float), out, st, cnt)
    paths = gio.write_regression_goldens(sim, str(tmp_path), "DAV2020-full")
    for mine, ref, tol in ((paths[0], "DAV2020-full_u_sol_referenceu_mm.csv", 1e-12), (paths[1], "DAV2020-full_u_sol_referenceu_δ.csv", 1e-6),
                           (paths[2], "DAV2020-full_u_sol_referenceu_belowground.csv", 1e-6)):
        a, b = gio.read_table(mine), gio.read_table(outputs(ref))
        assert list(a) == list(b), f"columns of {ref}"
        for k in b:
            assert np.array_equal(np.isnan(a[k]), np.isnan(b[k])), k
            scale = np.maximum(1.0, np.abs(b[k]))
            assert np.nanmax(np.abs(a[k] - b[k]) / scale, initial=0.0) < tol, k
    # flux table: header of the reference's flux goldens, ISO time stamps, exact round trip
    flux = gio.read_table(paths[3])
    ref_hdr = list(gio.read_table(outputs("INFEXP1_fluxes_referencedf.csv")))
    assert list(flux) == ref_hdr and len(flux["time"]) == 301
    assert flux["time"][0] == sim.parametrizedSPAC.reference_date and (flux["time"][1] - flux["time"][0]).days == 1
    direct = api.get_daily_soilFluxes(sim)
    assert all(np.array_equal(flux[k], np.asarray(direct[k], float), equal_nan=True) for k in ref_hdr[1:])
    # number formatting: shortest round-trip repr, NaN, signed zero, infinities
    t = {"time": np.array([0.0, 30.0, 60.0, 90.0]), "x": np.array([0.1, np.nan, -0.0, 1e-310]), "y": np.array([np.inf, -np.inf, 1 / 3, 5e22])}
    gio.write_table(str(tmp_path / "t.csv"), t)
    back = gio.read_table(str(tmp_path / "t.csv"))
    assert all(np.array_equal(back[k], t[k], equal_nan=True) for k in t) and np.signbit(back["x"][2])
    assert open(tmp_path / "t.csv").read().splitlines()[2] == "30.0,NaN,-Inf"


@pytest.mark.parametrize("which", ["dav_full", "bea", "isobea"])
def test_setup_twin_matches_batched_setup(which, request):
    """the member setup that runs on the GPU (include/lwfb90_setup.h, here compiled for the host with the
    device's elementary functions) against the numpy implementation pinned to remakeSPAC -> setup above: knob draw,
    soil-horizon scaling, field-capacity bisection, INFPAR, root profile, initial state.  Scalars are identical; the
    power-function results agree to the difference between two pow implementations."""
    import oracle as orc
    from lwfbrook90.jl_b200 import ensemble as en
    sim = request.getfixturevalue(which)
    M = 1500
    ens = en.make_ensemble(sim, M)
    kb = en.make_knob_block(sim, M)
    assert np.isnan(kb.knobs[0]).all(), "member 0 is the unmodified site definition: all knobs unset"
    scal, lay, u0, st = orc.setup_members(kb.N, kb.iso, kb.base_scal, kb.base_lay, kb.base_u0, kb.knobs, kb.over_idx,
                                          kb.profiles, kb.psim_base, kb.ilayer)
    assert (st == 0).all()
    assert np.array_equal(scal[0], sim.scal) and np.array_equal(lay[0], sim.lay) and np.array_equal(u0[0], sim.u0)

    def rel(a, b):
        sc = np.maximum(np.abs(a), np.abs(b)); sc[sc == 0] = 1
        return (np.abs(a - b) / sc).max()
    assert rel(scal, ens.scal) < 1e-13
    assert rel(lay, ens.lay) < 1e-10
    assert rel(u0, ens.u0) < 1e-13
    # an impossible soil (conductivity at saturation below the field-capacity conductivity): flagged, not silently used
    kb.knobs[5, en.K_KSAT_F] = 1e-9
    st = orc.setup_members(kb.N, kb.iso, kb.base_scal, kb.base_lay, kb.base_u0, kb.knobs, kb.over_idx, kb.profiles,
                           kb.psim_base, kb.ilayer)[3]
    assert st[5] == lo.ST.INPUT and (np.delete(st, 5) == 0).all()


def test_knob_rows_of_arbitrary_members(dav_full):
    """The synthetic ensemble is counter-based (Philox, counter = global member index): the knob rows of any list of members
    equal the rows of a contiguous block - what lets every rank of a multi-GPU run produce the members a cost-balanced deal
    gives it without moving data (bench.py --balance)."""
    from lwfbrook90.jl_b200 import ensemble as en
    block = en.make_knob_block(dav_full, 40)
    ids = np.array([17, 3, 39, 0, 8])
    picked = en.make_knob_block(dav_full, len(ids), members=ids)
    assert np.array_equal(picked.knobs, block.knobs[ids], equal_nan=True)
    shifted = en.make_knob_block(dav_full, 5, first_member=10)
    assert np.array_equal(shifted.knobs, block.knobs[10:15], equal_nan=True)
