"""Parity tests proper: the CUDA path, called through the C ABI (include/lwfb90.h), against the CPU
oracle on the same inputs.

Two comparisons:
  * oracle mode B (elementary functions of include/lwfb90_elem.h): the GPU record must be BIT-IDENTICAL
    for every member, every day, every state entry, including status and step counters;
  * oracle mode A (host libm, the build that is pinned to the reference's goldens): within the
    north-star tolerance (storage / theta 1e-6 relative, delta 1e-4 permil) wherever the reference
    algorithm itself is reproducible, which is established per member by re-running the oracle with a
    1e-15 relative perturbation of KSAT (see DESIGN.md "Parity").
"""
import os

import numpy as np
import pytest

from conftest import aligned_grid, inputs, outputs, read_table

import oracle as orc
from lwfbrook90.jl_b200 import api, capi, ensemble as en, layout as lo

pytestmark = pytest.mark.gpu

TOL_STORAGE = 1e-6      # relative, daily soil-water storage and layer theta (BASELINE.json north_star)
TOL_DELTA = 1e-4        # permil, delta18O / delta2H


def gpu_run(ens, n_days=None, members=None, record=True):
    sub = ens if members is None else ens.subset(members)
    nd = ens.n_days if n_days is None else n_days
    h = capi.Handle(ens.N, ens.iso, ens.t0_day, nd, record_daily=record)
    h.set_sites(ens.forcing)
    h.set_members(sub.site_of_member, sub.scal, sub.lay, sub.u0)
    h.run()
    return h, sub, nd


def oracle_run(sub, nd, mode, lay=None):
    orc.use(mode)
    return orc.run(sub.N, sub.iso, sub.t0_day, nd, sub.forcing, sub.site_of_member, sub.scal,
                   sub.lay if lay is None else lay, sub.u0, n_threads=8)


def assert_bit_identical(h, sub, nd):
    gpu = h.get_daily()
    st, cnt = h.get_status()
    ref, st_o, cnt_o = oracle_run(sub, nd, "elem")
    assert np.array_equal(st, st_o)
    assert np.array_equal(cnt, cnt_o), "accepted / rejected / RHS-evaluation counters differ"
    same = (gpu == ref) | (np.isnan(gpu) & np.isnan(ref))
    assert same.all(), f"{(~same).any(axis=(1, 2)).sum()} of {sub.M} members differ from the oracle"
    return gpu


# ------------------------------------------------------------------------------ building blocks
def test_elementary_functions_bit_equal_on_device():
    rng = np.random.default_rng(0)
    orc.use("elem")
    n = 4000
    cases = {0: rng.uniform(-700, 700, n), 1: 10 ** rng.uniform(-300, 300, n), 3: rng.uniform(-50, 50, n),
             4: rng.uniform(-50, 50, n), 5: rng.uniform(-1.5, 1.5, n), 6: rng.uniform(-1, 1, n), 7: rng.uniform(-1, 1, n)}
    for fn, x in cases.items():
        g, c = capi.elem_eval(fn, x), orc.elem(fn, x)
        assert np.array_equal(g, c, equal_nan=True), fn
    x, y = 10 ** rng.uniform(-6, 0, n), rng.uniform(-25, 25, n)
    assert np.array_equal(capi.elem_eval(2, x, y), orc.elem(2, x, y))
    # DiffEqBase-style fastpow of the step-size controller (Float32 arithmetic)
    x, y = 10 ** rng.uniform(-8, 3, n), np.where(rng.random(n) < 0.5, 0.14, 0.08)
    assert np.array_equal(capi.elem_eval(8, x, y), orc.elem(8, x, y))
    # edge cases: 0^y, x^0, log 0, exp of huge arguments, acos at the ends
    assert capi.elem_eval(2, [0.0, 0.0, 3.0, 0.0], [0.5, 0.0, 0.0, -1.0]).tolist() == [0.0, 1.0, 1.0, np.inf]
    assert capi.elem_eval(1, [0.0])[0] == -np.inf and capi.elem_eval(0, [-1e9, 1e9]).tolist() == [0.0, np.inf]
    assert np.isnan(capi.elem_eval(1, [-1.0])[0]) and capi.elem_eval(7, [1.0])[0] == 0.0


@pytest.mark.parametrize("which", ["dav_full", "bea", "infexp1"])
def test_rhs_bit_identical(which, request):
    """One evaluation of f(u,p,t) (lwfb_eval_rhs) for perturbed states, incl. clamped and saturated layers."""
    sim = request.getfixturevalue(which)
    ens = en.make_ensemble(sim, 64) if which != "infexp1" else en.batched_setup(sim, params={"DRAIN": np.linspace(0.1, 1, 64)})
    N, M = ens.N, ens.M
    rng = np.random.default_rng(5)
    ix = lo.StateIndex(N, ens.iso)
    smax = ens.lay[:, lo.L.SWATMAX]
    swati = ens.u0[:, ix.SWATI_mm] * rng.uniform(0.5, 1.3, (M, N))
    swati[::7] = smax[::7] * 1.01            # above saturation: SWCHEK clamps
    swati[3::11, 0] = -0.1                   # negative storage: clamped to 0
    gwat = rng.uniform(0, 5, M); gwat[::5] = 0.0
    frozen = np.concatenate([rng.uniform(0, 30, (M, 1)), rng.uniform(0, 2, (M, 1)), rng.uniform(0, 1, (M, 1)),
                             rng.uniform(-0.01, 0.5, (M, N))], axis=1)
    ntfli = rng.normal(0, 20, (M, N)); ntfli[::9] = np.nan      # first call of a fresh problem sees NaN
    h = capi.Handle(N, ens.iso, 0, 1)
    h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    orc.use("elem")
    for t in (0.0, 0.3712, 10.999999):
        du = h.eval_rhs(t, swati, gwat, frozen, ntfli)
        for m in range(M):
            ref, _ = orc.eval_rhs(N, ens.scal[m], ens.lay[m], t, swati[m], gwat[m], frozen[m], ntfli[m])
            assert np.array_equal(du[m], ref, equal_nan=True), (t, m)
    h.close()


# ------------------------------------------------------------------------------ full runs, bit-exact
def test_dav2020_full_single_member_vs_golden_G1(lw):
    """Member 0 = the reference's run_example(): GPU against the reference's own golden CSVs."""
    sim = lw.setup(lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True), requested_tspan=(0, 300))
    api.simulate_(sim)
    sol, ix = sim.solution, sim.solution.ix
    assert sol.retcode == "Success" and 24 * 300 <= sol.destats.naccept < 26 * 300
    h, a = read_table(outputs("DAV2020-full_u_sol_referenceu_belowground.csv"))
    days = a[:, 0]
    sw = np.array([sol.record_at(d)[ix.SWATI_mm] for d in days])
    assert np.max(np.abs(sw - a[:, 1:6]) / a[:, 1:6]) < TOL_STORAGE
    t = api.get_soil_(["θ", "ψ", "δ18O", "δ2H"], sim, days_to_read_out_d=days)
    th = np.stack([t[f"θ_Lay{i}"] for i in range(1, 6)], axis=1)
    assert np.max(np.abs(th - a[:, 16:21]) / a[:, 16:21]) < TOL_STORAGE
    d18 = np.stack([t[f"δ18O_Lay{i}"] for i in range(1, 6)], axis=1)
    d2 = np.stack([t[f"δ2H_Lay{i}"] for i in range(1, 6)], axis=1)
    assert np.max(np.abs(d18 - a[:, 6:11])) < TOL_DELTA and np.max(np.abs(d2 - a[:, 11:16])) < TOL_DELTA
    h, a = read_table(outputs("DAV2020-full_u_sol_referenceu_δ.csv"))
    d = api.get_delta(sim, days_to_read_out_d=a[:, 0])
    for j, name in enumerate(h[1:], start=1):
        assert np.array_equal(np.isnan(d[name]), np.isnan(a[:, j])), name
        assert np.nanmax(np.abs(d[name] - a[:, j])) < TOL_DELTA, name
    h, a = read_table(outputs("DAV2020-full_u_sol_referenceu_mm.csv"))
    am = api.get_amounts(sim, days_to_read_out_d=a[:, 0])
    for j, name in zip(range(1, 7), ["GWAT_mm", "INTS_mm", "INTR_mm", "SNOW_mm", "CC_MJm2", "SNOWLQ_mm"]):
        assert np.max(np.abs(am[name] - a[:, j])) <= 1e-9 + TOL_STORAGE * np.max(np.abs(a[:, j]))


@pytest.mark.parametrize("which,M,nd", [("dav_full", 96, None), ("bea", 48, None), ("isobea", 48, None)])
def test_ensemble_bit_identical_to_oracle(which, M, nd, request, monkeypatch):
    """Whole records of the three BASELINE site definitions, through both builds of the lane kernel: an ensemble that does
    not fill one block per SM runs in the latency build (one block per SM, 255 registers), LWFB_LATENCY_KERNEL=0 forces
    the throughput build (two blocks per SM, 128 registers) that large ensembles run in."""
    sim = request.getfixturevalue(which)
    ens = en.make_ensemble(sim, M)
    h, sub, nd = gpu_run(ens, nd)
    gpu = assert_bit_identical(h, sub, nd)
    # get_final returns the last record
    assert np.array_equal(h.get_final(), gpu[:, -1], equal_nan=True)
    monkeypatch.setenv("LWFB_LATENCY_KERNEL", "0")
    h.run()
    assert np.array_equal(h.get_daily(), gpu, equal_nan=True)
    h.close()


def test_generic_layer_count_kernels_bit_identical(lw, infexp1):
    """N = 42 (INFEXP1, saturation/bypass regime incl.) through the warp-per-member kernel (N > 32) and, forced, through
    the thread-per-member kernel; N = 13 (DAV2020-bare-minimum) through the lane kernel's general instantiation."""
    e = en.batched_setup(infexp1, params={"DRAIN": np.array([1.0, 0.6, 0.3])})
    h, sub, nd = gpu_run(e, 260)
    gpu = assert_bit_identical(h, sub, nd)
    h.force_legacy(True)
    h.run()
    assert np.array_equal(h.get_daily(), gpu, equal_nan=True)
    h.close()
    m = lw.loadSPAC(inputs("DAV2020-bare-minimum"), "DAV2020-minimal", simulate_isotopes=False, Δz_thickness_m=[0.1] * 11,
                    root_distribution=dict(beta=0.77, z_rootMax_m=-0.5), IC_soil=dict(PSIM_init_kPa=-7.0),
                    canopy_evolution=dict(DENSEF_rel=100, HEIGHT_rel=100, SAI_rel=100,
                                          LAI_rel=dict(DOY_Bstart=120, Bduration=21, DOY_Cstart=270, Cduration=60,
                                                       LAI_perc_BtoC=95, LAI_perc_CtoB=70)),
                    storm_durations_h=[5.44] * 12,
                    IC_scalar=dict(amount=dict(u_GWAT_init_mm=1., u_INTS_init_mm=13.7, u_INTR_init_mm=0., u_SNOW_init_mm=22.222,
                                               u_CC_init_MJ_per_m2=0.101010, u_SNOWLQ_init_mm=0.),
                                   d18O=dict(u_GWAT_init_permil=-11.111, u_INTS_init_permil=-12.222, u_INTR_init_permil=-13.333, u_SNOW_init_permil=-14.444),
                                   d2H=dict(u_GWAT_init_permil=-95.111, u_INTS_init_permil=-95.222, u_INTR_init_permil=-95.333, u_SNOW_init_permil=-95.444)))
    s = lw.setup(m)
    assert s.N == 13 and not s.iso
    e = en.make_ensemble(s, 8)
    h, sub, nd = gpu_run(e)
    assert_bit_identical(h, sub, nd)
    h.close()


@pytest.mark.parametrize("N,nd,M", [(40, 20, 40), (100, 4, 5), (400, 1, 3)])
def test_wide_kernel_deep_profiles(lw, N, nd, M):
    """The warp-per-member kernel (lwfb90_wide.cuh: 32 < N <= LWFB_MAX_LAYERS, per-layer arrays in shared memory, layers
    strided over the lanes) with isotopes on refined DAV2020-full grids up to the 400 layers of the reference's
    Hammel_sand-NLayer-400 integration test (test/02-integration-tests.jl:383): records, status, counters, du_NTFLI and
    the scores accumulated during the run, bit for bit against the oracle; one member aborts."""
    m = lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True, Δz_thickness_m=aligned_grid(N),
                    root_distribution=dict(beta=0.97, z_rootMax_m=None),
                    IC_soil=dict(PSIM_init_kPa=-6.0, delta18O_init_permil=-12.0, delta2H_init_permil=-90.0))
    s = lw.setup(m, ε=0.001, requested_tspan=(100, 100 + nd))
    assert s.N == N
    e = en.make_ensemble(s, M)
    e.u0[M - 1, N // 3] = np.nan                      # Unstable at its first step
    rng = np.random.default_rng(N)
    obs_day, obs_layer = np.arange(0, nd + 1), np.array([0, N // 2, N - 1])
    th, d18 = rng.uniform(0.1, 0.4, (nd + 1, 3)), rng.uniform(-14, -8, (nd + 1, 3))
    h = capi.Handle(e.N, e.iso, e.t0_day, nd)
    h.set_sites(e.forcing); h.set_members(e.site_of_member, e.scal, e.lay, e.u0)
    h.set_observations(obs_day, obs_layer, th, d18)
    h.run()
    gpu = assert_bit_identical(h, e, nd)
    st, _ = h.get_status()
    assert st[M - 1] == lo.ST.UNSTABLE and (st[:-1] == 0).all()
    from_sums = h.scores()
    nt = h.get_ntfli()
    orc.use("elem")
    o = orc.run_member(e.N, e.iso, e.t0_day, nd, e.forcing[0], e.scal[0], e.lay[0], e.u0[0])
    assert np.array_equal(nt[0], o[3])
    h2 = capi.Handle(e.N, e.iso, e.t0_day, nd)          # the record-based scorer on the same records
    h2.set_sites(e.forcing); h2.set_members(e.site_of_member, e.scal, e.lay, e.u0); h2.run()
    assert np.array_equal(from_sums, h2.scores(obs_day, obs_layer, th, d18), equal_nan=True)
    h.close(); h2.close()


def test_thread_per_member_kernel_agrees_with_lane_per_layer_kernel(dav_full, bea, isobea):
    """The library holds two independently written CUDA implementations (lane-per-layer production kernel,
    thread-per-member kernel for N < 3 or N > 32); both must reproduce the oracle bit-for-bit."""
    for sim, M in ((dav_full, 24), (bea, 12), (isobea, 40)):     # (isoBEA2016 member 22: miscompiled by NVVM -O3, see build.py)
        ens = en.make_ensemble(sim, M)
        h = capi.Handle(ens.N, ens.iso, ens.t0_day, 120)
        h.force_legacy(True)
        h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0); h.run()
        assert_bit_identical(h, ens, 120)
        h.close()


@pytest.mark.parametrize("N", [3, 4, 6, 8, 9, 16, 22, 32])
def test_lane_kernel_all_group_sizes(lw, N, monkeypatch):
    """Every lanes-per-member geometry (32/N members per warp, idle lanes, 1..3 extra components per lane), through both
    builds of the general instantiations (latency build: what a small ensemble runs in; throughput build forced)."""
    m = lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True, Δz_thickness_m=aligned_grid(N),
                    root_distribution=dict(beta=0.97, z_rootMax_m=None),
                    IC_soil=dict(PSIM_init_kPa=-6.0, delta18O_init_permil=-12.0, delta2H_init_permil=-90.0))
    s = lw.setup(m, ε=0.001 if N > 16 else 0.01, requested_tspan=(100, 130))
    assert s.N == N
    e = en.make_ensemble(s, 11)
    h, sub, nd = gpu_run(e)
    gpu = assert_bit_identical(h, sub, nd)
    monkeypatch.setenv("LWFB_LATENCY_KERNEL", "0")
    h.run()
    assert np.array_equal(h.get_daily(), gpu, equal_nan=True)
    h.close()


def test_adaptive_placement_is_result_neutral(dav_full):
    """Second and later runs of a handle place the members cost-sorted (lwfb_set_placement mode 1); records, status
    and counters must not change, with identity placement (mode 0) as the cross-check."""
    ens = en.make_ensemble(dav_full, 300)
    h, sub, nd = gpu_run(ens, 150)
    first = h.get_daily(); st1, cnt1 = h.get_status()
    h.set_placement(2)                        # mode 1 sorts only launches of several waves; 2 sorts always (and forgets
    h.run()                                   # the first run: this one is cold, in work quanta) ...
    h.run()                                   # ... and this one starts sorted by the previous run's RHS counts
    second = h.get_daily(); st2, cnt2 = h.get_status()
    assert np.array_equal(first, second, equal_nan=True) and np.array_equal(st1, st2) and np.array_equal(cnt1, cnt2)
    h.set_placement(0); h.run()
    assert np.array_equal(first, h.get_daily(), equal_nan=True)
    ref, st_o, cnt_o = oracle_run(sub, nd, "elem")
    assert np.array_equal(second, ref, equal_nan=True) and np.array_equal(cnt2, cnt_o)
    h.close()


@pytest.mark.parametrize("quantum", [60, 9, 2])
def test_work_quanta_with_regrouping_are_result_neutral(dav_full, monkeypatch, quantum):
    """lwfb_run integrates in work quanta: a launch ends after LWFB_QUANTUM passes of the block loop, every member
    checkpoints where it is (also mid-day), the unfinished members are regrouped by progress and resumed.  Records,
    status and counters must equal the oracle's single pass, for a cold handle and for a warm one (first launch
    cost-sorted)."""
    monkeypatch.setenv("LWFB_QUANTUM", str(quantum))
    if quantum == 9:
        monkeypatch.setenv("LWFB_SPREAD", "1")   # thin placement: one warp of members per block, the other slots empty
    ens = en.make_ensemble(dav_full, 200)
    nd = 60
    h = capi.Handle(ens.N, ens.iso, ens.t0_day, nd)
    h.set_placement(2)                        # quanta + regrouping regardless of the launch size
    h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    h.run()
    assert h.last_run_stats()[0] >= 60 * 20 // (16 * quantum)   # few blocks: quanta stretched up to 16 times
    assert_bit_identical(h, ens, nd)
    h.run()                                   # warm
    assert_bit_identical(h, ens, nd)
    h.close()


def test_work_quanta_priority_blocks_many_members(dav_full, monkeypatch):
    """More blocks than resident slots: the front blocks (least progress) keep their slots for a whole launch while the
    others turn over once per quantum; members of one ensemble end up with very different numbers of launches."""
    monkeypatch.setenv("LWFB_QUANTUM", "24")
    ens = en.make_ensemble(dav_full, 20000)
    nd = 6
    h = capi.Handle(ens.N, ens.iso, ens.t0_day, nd)
    h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    h.set_placement(2)
    h.run()
    assert h.last_run_stats()[0] >= 3
    assert_bit_identical(h, ens, nd)
    h.close()


@pytest.mark.parametrize("quantum", [0, 7])
def test_aborting_members(dav_full, monkeypatch, quantum):
    """retcode MaxIters (120 attempted steps) and Unstable (NaN in the initial state): records reached so far are kept,
    the rest is NaN, counters as the oracle's - in one launch and across work quanta."""
    monkeypatch.setenv("LWFB_QUANTUM", str(quantum))
    ens = en.make_ensemble(dav_full, 60)
    ens.u0[4, lo.StateIndex(ens.N, ens.iso).SWATI_mm.start + 2] = np.nan
    nd = 9
    h = capi.Handle(ens.N, ens.iso, ens.t0_day, nd, maxiters=120)
    h.set_placement(2)
    h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    h.run()
    gpu = h.get_daily(); st, cnt = h.get_status()
    orc.use("elem")
    ref, st_o, cnt_o = orc.run(ens.N, ens.iso, ens.t0_day, nd, ens.forcing, ens.site_of_member, ens.scal, ens.lay, ens.u0,
                               n_threads=8, maxiters=120)
    assert st_o[4] == 1 and (st_o == 3).sum() >= 50
    assert np.array_equal(st, st_o) and np.array_equal(cnt, cnt_o)
    assert np.array_equal(gpu, ref, equal_nan=True)
    h.close()


def test_nonzero_start_day_and_short_span(dav_full):
    """tspan not starting at 0 (initial conditions applied at t0, forcing taken from the right rows)."""
    e = en.make_ensemble(dav_full, 16)
    e.t0_day, e.n_days = 120, 45
    h, sub, nd = gpu_run(e)
    assert_bit_identical(h, sub, nd)
    h.close()


# ------------------------------------------------------------------------------ against the libm oracle
def test_tolerance_vs_libm_oracle_with_conditioning(dav_full, capsys):
    """GPU vs the golden-pinned oracle build (host libm, the reference's divisions) over the full record.

    The reference algorithm is chaotic for part of any ensemble: the north-star tolerance (1e-6 relative on storage and
    theta, 1e-4 permil on delta) is TIGHTER than the accuracy the solver is asked for (reltol 1e-5), so two correct
    implementations agree to it only while they take the same sequence of accepted and rejected steps, and a one-ulp
    difference in one pow() can flip one such decision.  What the test establishes:
      * the fraction of members whose whole record stays within the north-star tolerance is reported and must be >= 0.8
        (measured: 0.81 for these 64 members - 52 of 64; 0.83 with the round-1 exp - member 0, the unmodified site definition, is among them);
      * a member is CONDITIONED if the libm oracle itself stays within a tenth of the tolerance under six 1e-15 relative
        perturbations of its soil parameters; every conditioned member must be within tolerance;
      * every member outside the tolerance is one on which the oracle's own perturbed runs spread by a comparable amount
        (the GPU deviation is below 100 x that spread)."""
    ens = en.make_ensemble(dav_full, 64)
    h, sub, nd = gpu_run(ens)
    gpu = h.get_daily()
    h.close()
    ref = oracle_run(sub, nd, "libm")[0]
    ix = lo.StateIndex(sub.N, True)

    def errs(a, b):
        sw = np.abs(a[:, :, ix.SWATI_mm] - b[:, :, ix.SWATI_mm]) / np.abs(b[:, :, ix.SWATI_mm])
        st = np.abs(a[:, :, ix.acc("StorageSWAT")] - b[:, :, ix.acc("StorageSWAT")]) / b[:, :, ix.acc("StorageSWAT")]
        d = np.maximum(np.nan_to_num(np.abs(a[:, :, ix.SWATI_d18O] - b[:, :, ix.SWATI_d18O])),
                       np.nan_to_num(np.abs(a[:, :, ix.SWATI_d2H] - b[:, :, ix.SWATI_d2H])))
        return np.maximum(sw.max(axis=(1, 2)), st.max(axis=1)), d.max(axis=(1, 2))
    e_sw, e_d = errs(gpu, ref)
    s_sw, s_d = np.zeros(sub.M), np.zeros(sub.M)
    for field, eps in ((lo.L.KSAT, 1e-15), (lo.L.THSAT, 1e-15), (lo.L.ALPHA, 1e-15), (lo.L.KSAT, -1e-15), (lo.L.NPAR, 1e-15),
                       (lo.L.THICK, 1e-15)):
        lay_p = sub.lay.copy(); lay_p[:, field] *= (1 + eps)
        a, b = errs(oracle_run(sub, nd, "libm", lay=lay_p)[0], ref)
        s_sw, s_d = np.maximum(s_sw, a), np.maximum(s_d, b)
    within = (e_sw < TOL_STORAGE) & (e_d < TOL_DELTA)
    conditioned = (s_sw < TOL_STORAGE / 10) & (s_d < TOL_DELTA / 10)
    with capsys.disabled():
        print(f"\n[tolerance vs libm oracle] {sub.M} members x {nd} days: within the north-star tolerance {within.mean():.3f} "
              f"(storage/theta alone {np.mean(e_sw < TOL_STORAGE):.3f}); conditioned under perturbation {conditioned.mean():.3f}; "
              f"median deviation storage {np.median(e_sw):.1e}, delta {np.median(e_d):.1e} permil; "
              f"members outside: {np.nonzero(~within)[0].tolist()}")
    assert within[0] and conditioned[0], "the unmodified site definition must be well conditioned and within tolerance"
    assert within.mean() >= 0.8
    assert np.all(within[conditioned]), np.nonzero(conditioned & ~within)[0]
    assert np.median(e_sw) < 1e-9
    out = ~within
    assert np.all(e_sw[out] < 100 * np.maximum(s_sw[out], TOL_STORAGE)) and np.all(e_d[out] < 100 * np.maximum(s_d[out], TOL_DELTA))


# ------------------------------------------------------------------------------ size-independent properties
def test_properties_at_scale(dav_full):
    """20k members (beyond what the oracle finishes in seconds): determinism, permutation invariance,
    sub-sampled bit-exact check, physical invariants."""
    M = 20000
    ens = en.make_ensemble(dav_full, M)
    nd = 60
    h, _, _ = gpu_run(ens, nd)
    st, cnt = h.get_status()
    assert np.all(st == 0)
    fin = h.get_final()
    ix = lo.StateIndex(ens.N, True)
    # run again: bitwise deterministic
    h.run()
    assert np.array_equal(h.get_final(), fin, equal_nan=True)
    # permutation of members does not change any member's result (no cross-member coupling)
    perm = np.random.default_rng(1).permutation(M)
    h2, _, _ = gpu_run(ens, nd, members=perm)
    assert np.array_equal(h2.get_final(), fin[perm], equal_nan=True)
    h2.close()
    # physical invariants of the final state
    sw = fin[:, ix.SWATI_mm]
    assert np.all(sw >= 0) and np.all(sw <= ens.lay[:, lo.L.SWATMAX])
    assert np.all(fin[:, ix.GWAT_mm] >= 0) and np.all(fin[:, ix.SNOW_mm] >= 0) and np.all(fin[:, ix.INTS_mm] >= -1e-12)
    assert np.all(fin[:, ix.aux_theta] == 0) and np.all(fin[:, ix.XYLEM_mm] == 5.0)
    assert np.all(cnt[:, 0] >= 24 * nd) and np.all(cnt[:, 2] >= 6 * cnt[:, 0])
    # daily water balance of the last day closes unless SWCHEK clamped
    assert np.median(np.abs(fin[:, ix.acc("BALERD_SWAT")])) < 1e-9
    # a random sub-sample is bit-identical to the oracle
    pick = np.sort(np.random.default_rng(2).choice(M, 24, replace=False))
    sub = ens.subset(pick)
    ref, st_o, cnt_o = oracle_run(sub, nd, "elem")
    assert np.array_equal(np.stack([h.get_daily(int(i), 1)[0] for i in pick]), ref, equal_nan=True)
    assert np.array_equal(cnt[pick], cnt_o)
    h.close()


def test_bad_member_does_not_disturb_the_batch(dav_full):
    ens = en.make_ensemble(dav_full, 40)
    bad = ens.subset(np.arange(40))
    ix = lo.StateIndex(ens.N, True)
    bad.u0[5, ix.SWATI_mm.start + 1] = np.nan          # NaN state -> unstable_check
    bad.lay[9, lo.L.KSAT] = np.nan                      # NaN parameter
    h, _, nd = gpu_run(bad, 20)
    st, cnt = h.get_status()
    assert st[5] == 1 and st[9] == 1 and np.all(np.delete(st, [5, 9]) == 0)
    g = h.get_daily()
    h.close()
    h0, _, _ = gpu_run(ens, 20)
    g0 = h0.get_daily()
    h0.close()
    keep = np.delete(np.arange(40), [5, 9])
    assert np.array_equal(g[keep], g0[keep], equal_nan=True)
    assert np.all(np.isnan(g[5, 1:])) and np.all(np.isnan(g[9, 1:]))     # records never reached are NaN
    # the oracle reports the same status words
    ref, st_o, _ = oracle_run(bad, 20, "elem")
    assert np.array_equal(st, st_o)


def test_scores_against_numpy(isobea):
    ens = en.make_ensemble(isobea, 32)
    nd = 90
    h, sub, _ = gpu_run(ens, nd)
    rec = h.get_daily()
    ix = lo.StateIndex(ens.N, True)
    obs_day = np.arange(5, nd, 7).astype(np.int32)
    obs_layer = np.array([0, 2, 4], dtype=np.int32)
    # synthetic observations = member 3 + noise, with gaps
    rng = np.random.default_rng(7)
    sw = rec[3][obs_day][:, obs_layer]
    ths, thr, smax = (ens.lay[3, k][obs_layer] for k in (lo.L.THSAT, lo.L.THR, lo.L.SWATMAX))
    th_obs = np.minimum(1.0, (ths * sw / smax - thr) / (ths - thr)) * (ths - thr) + thr + rng.normal(0, 0.005, sw.shape)
    d_obs = rec[3][obs_day][:, ix.SWATI_d18O][:, obs_layer] + rng.normal(0, 0.1, sw.shape)
    th_obs[2, 1] = np.nan; d_obs[4, 0] = np.nan
    sc = h.scores(obs_day, obs_layer, th_obs, d_obs)
    for m in range(ens.M):
        swm = rec[m][obs_day][:, obs_layer]
        thsm, thrm, smaxm = (ens.lay[m, k][obs_layer] for k in (lo.L.THSAT, lo.L.THR, lo.L.SWATMAX))
        sim_th = np.minimum(1.0, (thsm * swm / smaxm - thrm) / (thsm - thrm)) * (thsm - thrm) + thrm
        sim_d = rec[m][obs_day][:, ix.SWATI_d18O][:, obs_layer]
        for v, (sim, obs) in enumerate(((sim_th, th_obs), (sim_d, d_obs))):
            ok = ~np.isnan(obs)
            s_, o_ = sim[ok], obs[ok]
            rmse = np.sqrt(np.mean((s_ - o_) ** 2))
            r = np.corrcoef(s_, o_)[0, 1]
            kge = 1 - np.sqrt((r - 1) ** 2 + (s_.std() / o_.std() - 1) ** 2 + (s_.mean() / o_.mean() - 1) ** 2)
            assert np.isclose(sc[m, 2 * v], rmse, rtol=1e-9) and np.isclose(sc[m, 2 * v + 1], kge, rtol=1e-6, atol=1e-9)
    assert np.argmin(sc[:, 0]) == 3 and np.argmin(sc[:, 2]) == 3     # the generating member scores best
    h.close()


def test_abi_error_handling(dav_full):
    ens = en.make_ensemble(dav_full, 4)
    h = capi.Handle(ens.N, ens.iso, 0, 10)
    with pytest.raises(capi.LwfbError, match="set_sites"):
        h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    with pytest.raises(capi.LwfbError, match="shorter"):
        h.set_sites(ens.forcing[:, :5])
    h.set_sites(ens.forcing)
    with pytest.raises(capi.LwfbError, match="site index"):
        h.set_members(np.full(4, 3), ens.scal, ens.lay, ens.u0)
    h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    with pytest.raises(capi.LwfbError):
        h.get_daily()                               # not run yet
    h.run()
    with pytest.raises(capi.LwfbError):
        h.get_daily(2, 10)                          # out of range
    h.close()
    h = capi.Handle(ens.N, ens.iso, 0, 10, record_daily=False)
    h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0); h.run()
    with pytest.raises(capi.LwfbError, match="record_daily"):
        h.get_daily()
    assert h.get_final().shape == (4, ens.u0.shape[1])
    h.close()


def water_partitioning_numpy(rec, ix, group, n_groups):
    """Sequential-order restatement of get_water_partitioning (src/func_postprocess.jl:1310-1364) for a block of members:
    rec [M, n_rec, L] -> [M, n_groups, NWP].  np.cumsum adds in record order, like the device kernel."""
    A = lambda n: rec[:, :, ix.acc(n)]
    daily = np.empty(rec.shape[:2] + (lo.NWP,))
    daily[..., lo.WP.ETA] = A("cum_d_evap"); daily[..., lo.WP.ESOIL] = A("cum_d_slvp"); daily[..., lo.WP.ESNOW] = A("cum_d_snvp")
    daily[..., lo.WP.EINTERCEPTION] = A("cum_d_irvp") + A("cum_d_isvp")
    daily[..., lo.WP.TA] = A("cum_d_tran"); daily[..., lo.WP.PRECIP] = A("cum_d_prec")
    daily[..., lo.WP.TD] = A("cum_d_ptran") - A("cum_d_tran")
    daily[..., lo.WP.D] = -A("vrfln"); daily[..., lo.WP.R] = -(A("srfl") + A("byfl") + A("dsfl"))
    daily[..., lo.WP.SWAT] = A("StorageSWAT")
    out = np.full((rec.shape[0], n_groups, lo.NWP), np.nan)
    for g in range(n_groups):
        sel = np.flatnonzero(group == g)
        if sel.size == 0:
            out[:, g, :lo.WP.SWAT] = 0.0
            continue
        out[:, g] = np.cumsum(daily[:, sel], axis=1)[:, -1]
        out[:, g, lo.WP.SWAT] /= sel.size
    return out


def test_water_partitioning_on_device(dav_full):
    """N3 of SURVEY 8f: monthly / yearly water partitioning of every member aggregated on the GPU from the resident
    daily record — bit-identical to the sequential numpy restatement, and equal (1e-12) to the reference-facing
    get_water_partitioning of the single-member API on member 0."""
    import datetime as dt
    ens = en.make_ensemble(dav_full, 70)
    nd = 100
    res = api.simulate_ensemble(ens, n_days=nd)
    rec = res.daily()
    ix = lo.StateIndex(ens.N, ens.iso)
    ref_date = dav_full.parametrizedSPAC.reference_date
    for period in ("month", "year"):
        labels, tab = res.water_partitioning(ref_date, period)
        dates = [ref_date + dt.timedelta(days=float(t)) for t in range(ens.t0_day, ens.t0_day + nd + 1)]
        lab = [(d.year, d.month) if period == "month" else (d.year,) for d in dates]
        group = np.array([labels.index(l) for l in lab], dtype=np.int32)
        assert tab.shape == (ens.M, len(labels), lo.NWP)
        assert np.array_equal(tab, water_partitioning_numpy(rec, ix, group, len(labels)))
    # records that are not used (group < 0) and an empty period
    group = np.full(nd + 1, -1, dtype=np.int32); group[10:40] = 0; group[40:41] = 2
    tab = res.handle.water_partitioning(group, 3)
    want = water_partitioning_numpy(rec, ix, group, 3)
    assert np.array_equal(tab, want, equal_nan=True) and np.isnan(tab[:, 1, lo.WP.SWAT]).all()
    # against the reference-facing single-member function (numpy pairwise sums): member 0 = the unmodified site
    import copy
    sim = copy.deepcopy(dav_full)             # fresh object: simulate_ stores its solution in it
    api.simulate_(sim, tspan=(ens.t0_day, ens.t0_day + nd))
    _, monthly, _ = api.get_water_partitioning(sim)
    labels, tab = res.water_partitioning(ref_date, "month")
    for c, name in ((lo.WP.ETA, "ETa"), (lo.WP.TD, "Td"), (lo.WP.D, "D"), (lo.WP.R, "R"), (lo.WP.PRECIP, "Precip"), (lo.WP.SWAT, "Swat")):
        np.testing.assert_allclose(tab[0, :, c], monthly[name], rtol=1e-12, atol=1e-12)
    res.close()


def test_soil_state_on_device(dav_full, bea):
    """N3 of SURVEY 8f: get_soil_ for every member on the GPU.  theta, W, SWATI and the deltas are bit-identical to the
    host post-processing (same IEEE operations); psi and K go through pow and agree to 1e-9 with the numpy twin of
    derive_auxiliary_SOILVAR (src/module_KPT.jl:432-443)."""
    from lwfbrook90.jl_b200 import kpt
    for sim in (dav_full, bea):
        ens = en.make_ensemble(sim, 40)
        nd = 30
        res = api.simulate_ensemble(ens, n_days=nd)
        rec = res.daily()
        ix = lo.StateIndex(ens.N, ens.iso)
        days = np.array([0, 1, 7, 30])
        layers = np.array([0, ens.N - 1, 2])
        got = res.soil_state(days, layers=layers)
        assert got.shape == (ens.M, 4, 3, lo.NSV)
        S = rec[:, days][:, :, ix.SWATI_mm][:, :, layers]
        P = lambda k: ens.lay[:, k][:, None, layers]
        W = np.minimum(1.0, (P(lo.L.THSAT) * S / P(lo.L.SWATMAX) - P(lo.L.THR)) / (P(lo.L.THSAT) - P(lo.L.THR)))
        assert np.array_equal(got[..., lo.SV.SWATI], S) and np.array_equal(got[..., lo.SV.W], W)
        assert np.array_equal(got[..., lo.SV.THETA], kpt.FTheta_MvG(W, P(lo.L.THSAT), P(lo.L.THR)))
        np.testing.assert_allclose(got[..., lo.SV.PSI], kpt.FPSIM_MvG(W, P(lo.L.ALPHA), P(lo.L.NPAR), P(lo.L.MPAR)), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(got[..., lo.SV.K], kpt.FK_MvG(W, P(lo.L.KSAT), P(lo.L.TORT), P(lo.L.NPAR)), rtol=1e-7)   # 1 - (1 - A^..)^.. cancels in both implementations
        if ens.iso:
            assert np.array_equal(got[..., lo.SV.D18O], rec[:, days][:, :, ix.SWATI_d18O][:, :, layers], equal_nan=True)
            assert np.array_equal(got[..., lo.SV.D2H], rec[:, days][:, :, ix.SWATI_d2H][:, :, layers], equal_nan=True)
        else:
            assert np.isnan(got[..., lo.SV.D18O]).all()
        # depth lookup as get_soil_idx: 100 mm and 500 mm of the shared grid
        by_depth = res.soil_state(days, depths_to_read_out_mm=[100, 500])
        idx = api.get_soil_idx(sim, [100, 500], only_valid_idxs=True)
        assert np.array_equal(by_depth, res.soil_state(days, layers=[i - 1 for i in idx]), equal_nan=True)
        res.close()


def test_scores_accumulated_by_the_integrator(isobea, monkeypatch):
    """lwfb_set_observations: scores without the resident daily record (record_daily = 0) are bit-identical to the
    scores formed from the record, in a single launch and in work quanta, with a member that aborts."""
    ens = en.make_ensemble(isobea, 40)
    ens.u0[7, 3] = np.nan                     # member 7: NaN in SWATI -> Unstable
    nd = 60
    obs_day = np.array([0, 3, 10, 11, 30, 59, 60], dtype=np.int32)
    obs_layer = np.array([0, 3, 6], dtype=np.int32)
    rng = np.random.default_rng(11)
    th = 0.25 + 0.05 * rng.random((len(obs_day), 3)); d18 = -10 + rng.random((len(obs_day), 3))
    th[2, 1] = np.nan; d18[5, 2] = np.nan
    # (a) from the record
    h, sub, _ = gpu_run(ens, nd)
    want = h.scores(obs_day, obs_layer, th, d18)
    st_a, _ = h.get_status()
    h.close()
    assert st_a[7] == lo.ST.UNSTABLE and np.isnan(want[7]).all() and np.isfinite(np.delete(want, 7, axis=0)).all()
    # (b) accumulated during the run, no record; single launch and quanta of 5 passes with regrouping
    for quantum in (0, 5):
        monkeypatch.setenv("LWFB_QUANTUM", str(quantum))
        h = capi.Handle(ens.N, ens.iso, ens.t0_day, nd, record_daily=False)
        h.set_placement(2)
        h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
        with pytest.raises(capi.LwfbError, match="lwfb_set_observations"):
            h.run(); h.scores()
        with pytest.raises(capi.LwfbError, match="record_daily"):
            h.scores(obs_day, obs_layer, th, d18)
        h.set_observations(obs_day, obs_layer, th, d18)
        h.run()
        got = h.scores()
        assert np.array_equal(got, want, equal_nan=True), quantum
        h.run()                               # a second run starts its sums afresh
        assert np.array_equal(h.scores(), want, equal_nan=True)
        h.close()
    # the thread-per-member kernel accumulates the same sums
    h = capi.Handle(ens.N, ens.iso, ens.t0_day, nd, record_daily=False)
    h.force_legacy(True)
    h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    h.set_observations(obs_day, obs_layer, th, d18)
    h.run()
    assert np.array_equal(h.scores(), want, equal_nan=True)
    h.close()


def test_getters_wait_for_an_asynchronous_run(dav_full):
    """lwfb_run_async on a caller's non-blocking stream followed directly by a getter: the getter synchronises."""
    import torch
    ens = en.make_ensemble(dav_full, 600)
    nd = 40
    h, sub, _ = gpu_run(ens, nd)
    want = h.get_daily(); st_w, cnt_w = h.get_status()
    h.close()
    h = capi.Handle(ens.N, ens.iso, ens.t0_day, nd)
    h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    s = torch.cuda.Stream()
    h.run_async(s.cuda_stream)
    st, cnt = h.get_status()                  # no lwfb_sync in between
    assert np.array_equal(st, st_w) and np.array_equal(cnt, cnt_w)
    h.run_async(s.cuda_stream)
    got = h.get_daily()
    assert np.array_equal(got, want, equal_nan=True)
    # a pinned destination takes the direct path, a large pageable one the staged path: same bytes
    pinned = torch.empty(want.shape, dtype=torch.float64).pin_memory()
    h.get_daily(out=pinned.numpy())
    assert np.array_equal(pinned.numpy(), want, equal_nan=True)
    with pytest.raises(capi.LwfbError, match="get_daily"):
        h.get_daily(out=np.empty((3, 3)))
    h.close()


@pytest.mark.parametrize("which", ["dav_full", "bea"])
def test_setup_members_on_device(which, request):
    """SURVEY 8f N1 on the GPU: lwfb_setup_members builds the member blocks in HBM from the base member and a knob row per
    member; they must be bit-identical to what the host compile of the same text (oracle twin) produces, for 10^5
    members, and a run from them bit-identical to a run from the uploaded blocks."""
    import time
    sim = request.getfixturevalue(which)
    M = 100_000
    kb = en.make_knob_block(sim, M)
    kb.knobs[77, en.K_KSAT_F] = 1e-9          # an impossible soil: reported, does not disturb the others
    want = orc.setup_members(kb.N, kb.iso, kb.base_scal, kb.base_lay, kb.base_u0, kb.knobs, kb.over_idx, kb.profiles,
                             kb.psim_base, kb.ilayer)
    h = capi.Handle(kb.N, kb.iso, kb.t0_day, 5, record_daily=False)
    h.set_sites(kb.forcing)
    h.setup_members(kb)                       # warm-up (allocations)
    t = time.perf_counter()
    n_invalid = h.setup_members(kb)
    dt = time.perf_counter() - t
    assert n_invalid == int((want[3] != 0).sum()) == 1
    for k in range(3):
        assert np.array_equal(h.get_block(k), want[k], equal_nan=True), ("scal", "lay", "u0")[k]
    print(f"\n[setup_members] {M} members in {dt * 1e3:.1f} ms (H2D of the knob rows included)")
    assert dt < 0.25
    # integrate a few days from the device-built blocks and from the same blocks uploaded by lwfb_set_members
    sub = np.r_[0:64, 99_936:100_000]
    kb2 = en.KnobBlock(**{**kb.__dict__, "knobs": kb.knobs[sub]})
    h.setup_members(kb2); h.run()
    fin_a = h.get_final(); st_a, cnt_a = h.get_status()
    h.set_members(np.zeros(len(sub), np.int32), want[0][sub], want[1][sub], want[2][sub]); h.run()
    assert np.array_equal(fin_a, h.get_final(), equal_nan=True) and np.array_equal(cnt_a, h.get_status()[1])
    h.close()


# ------------------------------------------------------------------------------ the reference's goldens, straight from the GPU
def _gpu_single(sim, n_days=None, ntfli0=None, legacy=False):
    nd = int(sim.tspan[1] - sim.tspan[0]) if n_days is None else n_days
    h = capi.Handle(sim.N, sim.iso, int(sim.tspan[0]), nd)
    if legacy:
        h.force_legacy(True)
    h.set_sites(sim.forcing); h.set_members([0], sim.scal, sim.lay, sim.u0)
    if ntfli0 is not None:
        h.set_ntfli0(ntfli0)
    h.run()
    rec = h.get_daily()[0]
    st, cnt = h.get_status()
    nt = h.get_ntfli()
    h.close()
    assert st[0] == 0
    return rec, cnt[0], nt


@pytest.mark.parametrize("modified", [False, True])
def test_G2_G3_infexp1_golden_through_the_gpu(lw, infexp1, modified):
    """Goldens G2 / G3 (examples/INFEXP1*_fluxes_referencedf.csv, run definition test/03-regression-tests.jl:352-367)
    compared DIRECTLY with the GPU record (N = 42: warp-per-member kernel), with the bounds the oracle is held to."""
    from test_oracle_golden import _flux_cmp
    sim = infexp1
    if modified:
        sim = lw.remakeSPAC(sim, params=dict(BYPAR=1, QDEPTH_m=0.40, QFFC=0.2, QFPAR=0.3, GSC=0.05, GSP=0.2))
    assert sim.N == 42
    rec, _, _ = _gpu_single(sim)
    name = "INFEXP1-modified_fluxes_referencedf.csv" if modified else "INFEXP1_fluxes_referencedf.csv"
    hdr, mine, ref = _flux_cmp(rec, lo.StateIndex(42, False), outputs(name))
    assert np.max(np.abs(mine[:232] - ref[:232])) < 1e-7          # well conditioned until the profile saturates (day 231)
    for c in ("StorageSWAT", "StorageWATER"):
        j = hdr.index(c)
        assert np.max(np.abs(mine[:, j] - ref[:, j]) / ref[:, j]) < 2e-3
    for c in ("cum_d_prec", "cum_d_rfal", "cum_d_sfal", "cum_d_rint", "cum_d_pint", "cum_d_irvp"):
        j = hdr.index(c)
        assert np.allclose(mine[:, j], ref[:, j], rtol=1e-4, atol=1e-4)


def test_G4_dav2020_modified_golden_through_the_gpu(lw):
    """Golden G4 (examples/DAV2020-full-modified_fluxes_referencedf.csv, test/03-regression-tests.jl:411-458; N = 40,
    every optional flow path switched on: warp-per-member kernel) against the GPU record at the reference's own tolerance,
    all 30 series."""
    from test_oracle_golden import _dav_modified, _flux_cmp
    sim = _dav_modified(lw)
    assert sim.N == 40
    rec, _, _ = _gpu_single(sim)
    _, mine, ref = _flux_cmp(rec, lo.StateIndex(40, True), outputs("DAV2020-full-modified_fluxes_referencedf.csv"))
    assert np.all(np.abs(mine - ref) <= 1e-4 + 1e-4 * np.abs(ref))
    assert np.max(np.abs(mine - ref)) < 1e-5


def test_G5_second_simulate_on_the_gpu(lw):
    """Golden G5 (inline tables of test/01-unit-tests.jl:907-1014) comes from the SECOND simulate! on one problem: the
    first ITER sees the du_NTFLI the first run left in `p`.  lwfb_get_ntfli / lwfb_set_ntfli0 reproduce
    that on the GPU (N = 22: the lane kernel's general instantiation); isotope table and water partitioning against the
    golden, and the warm run bit for bit against the oracle's."""
    import datetime as dt
    import json
    import os
    from test_oracle_golden import GOLD, _g5_model
    sim = _g5_model(lw)
    cold, _, nt = _gpu_single(sim)
    warm, cnt, _ = _gpu_single(sim, ntfli0=nt)
    orc.use("elem")
    o_cold = orc.run_member(sim.N, sim.iso, 0, 365, sim.forcing, sim.scal, sim.lay, sim.u0)
    o_warm = orc.run_member(sim.N, sim.iso, 0, 365, sim.forcing, sim.scal, sim.lay, sim.u0, ntfli_init=o_cold[3])
    assert np.array_equal(nt[0], o_cold[3]) and np.array_equal(warm, o_warm[0], equal_nan=True) and np.array_equal(cnt, o_warm[2])
    gold = json.load(open(os.path.join(GOLD, "golden_inline.json")))
    sim.solution = api.ODESolution(sim, np.arange(366.0), warm, 0, cnt)
    sim.ODESolution_datetime = [sim.parametrizedSPAC.reference_date + dt.timedelta(days=k) for k in range(366)]
    ref = np.array(gold["reference_isotopePlotPermutedDims"])
    sol, ix = sim.solution, sim.solution.ix
    days = [0, 9, 99, 299]
    rec = np.array([sol.record_at(d) for d in days])
    for j, n in enumerate(["INTS", "INTR", "SNOW", "GWAT", "RWU", "XYLEM"]):
        assert np.allclose(rec[:, getattr(ix, n + "_d18O")], ref[6 + j], rtol=0, atol=1e-6)
        assert np.allclose(rec[:, getattr(ix, n + "_d2H")], ref[12 + j], rtol=0, atol=1e-5)
    assert np.max(np.abs(rec[:, ix.SWATI_d18O].T - ref[19:41])) < 1e-6
    assert np.max(np.abs(rec[:, ix.SWATI_d2H].T - ref[41:63])) < 1e-5
    daily, monthly, yearly = api.get_water_partitioning(sim)
    cols = ["ETa", "Esoil", "Esnow", "Einterception", "Ta", "Precip", "Td", "D", "R", "Swat"]
    for tab, key, rows in ((daily, "reference_daily_check", [9, 19, 99, 179, 249, 259]), (monthly, "reference_monthly_check", None),
                           (yearly, "reference_yearly_check", None)):
        rows = range(len(tab["year"])) if rows is None else rows
        mine, want = np.array([[tab[c][r] for c in cols] for r in rows]), np.array(gold[key])
        assert np.linalg.norm(mine - want) <= max(1e-5, 1e-5 * max(np.linalg.norm(mine), np.linalg.norm(want)))


def test_rhs_by_both_kernels(dav_full):
    """lwfb_eval_rhs runs the kernel that would integrate the members: the lane kernel (lean and general
    instantiation) and, forced, the thread-per-member kernel give the same bits as the oracle."""
    ens = en.make_ensemble(dav_full, 40)
    N, M = ens.N, ens.M
    rng = np.random.default_rng(9)
    ix = lo.StateIndex(N, ens.iso)
    swati = ens.u0[:, ix.SWATI_mm] * rng.uniform(0.5, 1.3, (M, N))
    gwat = rng.uniform(0, 5, M)
    frozen = np.concatenate([rng.uniform(0, 30, (M, 1)), rng.uniform(0, 2, (M, 1)), rng.uniform(0, 1, (M, 1)),
                             rng.uniform(-0.01, 0.5, (M, N))], axis=1)
    ntfli = rng.normal(0, 20, (M, N)); ntfli[::9] = np.nan
    orc.use("elem")
    want = np.stack([orc.eval_rhs(N, ens.scal[m], ens.lay[m], 3.25, swati[m], gwat[m], frozen[m], ntfli[m])[0] for m in range(M)])
    h = capi.Handle(N, ens.iso, 0, 1)
    h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    for mode in (0, 2, 1):                      # lean lane kernel, general lane kernel, thread-per-member kernel
        h.force_legacy(mode)
        assert np.array_equal(h.eval_rhs(3.25, swati, gwat, frozen, ntfli), want, equal_nan=True), mode
    h.close()


def test_simulate_keyword_arguments(dav_full):
    """simulate_ forwards the solver keywords it supports and refuses what it cannot honour."""
    import copy
    sim = copy.copy(dav_full)
    with pytest.raises(NotImplementedError, match="save_everystep"):
        api.simulate_(sim, save_everystep=True)
    with pytest.raises(NotImplementedError, match="day boundaries"):
        api.simulate_(sim, saveat=[0.0, 1.5])
    with pytest.raises(TypeError, match="unsupported"):
        api.simulate_(sim, alg="Rodas5")
    api.simulate_(sim, tspan=(0, 20), saveat=[0, 10, 20])
    base = sim.solution.records.copy()
    assert list(sim.solution.t) == [0.0, 10.0, 20.0]
    api.simulate_(sim, tspan=(0, 20), saveat=[0, 10, 20], reltol=1e-7)
    assert sim.solution.destats.naccept > 0 and not np.array_equal(sim.solution.records, base)
    assert np.allclose(sim.solution.records[:, :5], base[:, :5], rtol=1e-4)


@pytest.mark.parametrize("balanced", [False, True])
def test_multi_handle_shards_and_allgathers(dav_full, balanced):
    """lwfb_multi_* (several GPUs of one host behind the C ABI, no torch / NCCL): shards dealt contiguously or balanced
    by a cost estimate, run concurrently, scores all-gathered peer to peer - everything in the caller's member order and
    bit-identical to one handle.  Uses every visible GPU, and at least two shards (two on one device if need be)."""
    import torch
    n_dev = torch.cuda.device_count()
    devices = list(range(n_dev)) if n_dev >= 2 else [0, 0]
    ens = en.make_ensemble(dav_full, 301)
    nd = 25
    obs_day = np.array([3, 10, 24], dtype=np.int32); obs_layer = np.array([0, 2], dtype=np.int32)
    th = np.full((3, 2), 0.21); d18 = np.full((3, 2), -11.0)
    h = capi.Handle(ens.N, ens.iso, 0, nd, record_daily=False)
    h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    h.set_observations(obs_day, obs_layer, th, d18)
    h.run()
    want_sc, want_fin = h.scores(), h.get_final()
    want_st, want_cnt = h.get_status()
    h.close()
    mh = capi.MultiHandle(ens.N, ens.iso, 0, nd, devices)
    mh.set_sites(ens.forcing)
    cost = want_cnt[:, 2].astype(float) if balanced else None
    mh.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0, cost=cost)
    mh.set_observations(obs_day, obs_layer, th, d18)
    mh.run()
    sc, ptrs = mh.allgather_scores()
    st, cnt, shard = mh.get_status()
    assert np.array_equal(sc, want_sc, equal_nan=True) and np.array_equal(st, want_st) and np.array_equal(cnt, want_cnt)
    assert np.array_equal(mh.get_final(), want_fin, equal_nan=True)
    sizes = np.bincount(shard, minlength=len(devices))
    assert sizes.max() - sizes.min() <= 1 and len(set(ptrs)) == len(devices)
    if balanced:      # every shard carries the same share of the cost (snake deal of the sorted members)
        load = np.array([cost[shard == s].sum() for s in range(len(devices))])
        assert load.max() / load.min() < 1.02
    else:
        assert (np.diff(shard) >= 0).all()
    # the gathered block on every device equals the host copy
    for s, p in enumerate(ptrs):
        from lwfbrook90.jl_b200.parallel import DeviceArray
        t = torch.as_tensor(DeviceArray(p, sc.shape), device=f"cuda:{devices[s]}")
        assert np.array_equal(t.cpu().numpy(), sc, equal_nan=True)
    mh.close()


def test_c_host_drives_the_multi_gpu_abi(isobea, tmp_path):
    """A C program (tests/c/multi_host.c, compiled with gcc against include/lwfb90.h; no Python, torch or NCCL in its
    process) shards an ensemble over lwfb_multi_* with a cost-balanced deal, integrates, all-gathers the scores peer to
    peer and returns everything in the caller's member order: bit-equal to a single handle driven from Python.  Uses two
    shards (both GPUs if the box has two, else two shards on device 0)."""
    import subprocess
    import torch
    ens = en.make_ensemble(isobea, 70)
    nd = 40
    N, L, M = ens.N, ens.u0.shape[1], ens.M
    obs_day = np.array([0, 5, 17, 40], dtype=np.int32); obs_layer = np.array([0, 3, 6], dtype=np.int32)
    rng = np.random.default_rng(21)
    th = 0.25 + 0.05 * rng.random((4, 3)); d18 = -10 + rng.random((4, 3))
    cost = rng.random(M)
    src, out = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(src, "wb") as f:
        def put(a, dt):
            b = np.ascontiguousarray(a, dtype=dt).tobytes()
            f.write(b + b"\0" * (-len(b) % 8))
        put([N, int(ens.iso), ens.t0_day, nd, ens.forcing.shape[-2], len(obs_day), len(obs_layer), 0], np.int32)
        put([M], np.int64)
        put(ens.forcing.reshape(-1, ens.forcing.shape[-1]), np.float64); put(ens.site_of_member, np.int32)
        put(ens.scal, np.float64); put(ens.lay, np.float64); put(ens.u0, np.float64); put(cost, np.float64)
        put(obs_day, np.int32); put(obs_layer, np.int32); put(th, np.float64); put(d18, np.float64)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(capi.LIB_PATH)
    exe = tmp_path / "multi_host"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-O1", "-o", str(exe), os.path.join(root, "tests", "c", "multi_host.c"),
                           "-L", libdir, "-l:" + os.path.basename(capi.LIB_PATH), "-Wl,-rpath," + libdir])
    devices = ["0", "1"] if torch.cuda.device_count() >= 2 else ["0", "0"]
    r = subprocess.run([str(exe), str(src), str(out)] + devices, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    o = 0
    def take(n, dt):
        nonlocal o
        a = np.frombuffer(raw, dtype=dt, count=n, offset=o); o += a.nbytes
        return a
    fin = take(M * L, np.float64).reshape(M, L); scores = take(M * 4, np.float64).reshape(M, 4)
    counters = take(M * 3, np.int64).reshape(M, 3); status = take(M, np.int32); shard = take(M, np.int32)
    assert sorted(np.bincount(shard, minlength=2)) == [35, 35]
    h = capi.Handle(N, ens.iso, ens.t0_day, nd, record_daily=False)
    h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0)
    h.set_observations(obs_day, obs_layer, th, d18)
    h.run()
    st, cnt = h.get_status()
    assert np.array_equal(status, st) and np.array_equal(counters, cnt)
    assert np.array_equal(fin, h.get_final(), equal_nan=True) and np.array_equal(scores, h.scores(), equal_nan=True)
    h.close()
