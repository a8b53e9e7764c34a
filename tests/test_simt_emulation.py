"""Kernel text under the lock-step warp emulator (tests/simt) against the CPU oracle, bit for bit.

These tests need no GPU: the CUDA kernel headers of lwfbrook90.jl_b200/csrc are compiled for the host with
-DLWFB_SIMT_EMUL and executed with SIMT semantics (the threads of a block as coroutines that meet at every warp
primitive and block barrier; threads meeting at different primitives abort).  They check the arithmetic, the shuffle
choreography and the control flow of the kernels — what the GPU parity tests (-m gpu) check on the
device — on the machine that builds the repository.  The emulator is test infrastructure only.
"""
import os
import sys

import numpy as np
import pytest

from conftest import aligned_grid, inputs

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
import emul  # noqa: E402
import oracle as orc  # noqa: E402
from lwfbrook90.jl_b200 import ensemble as en, layout as lo  # noqa: E402


def check(ens, n_days, kernel, members=None):
    sub = ens if members is None else ens.subset(np.arange(members))
    orc.use("elem")
    ref, st_o, cnt_o = orc.run(sub.N, sub.iso, sub.t0_day, n_days, sub.forcing, sub.site_of_member, sub.scal, sub.lay,
                               sub.u0, n_threads=8)
    emul.use_wide(kernel == emul.WIDE)
    try:
        got, st, cnt = emul.run(kernel, sub.N, sub.iso, sub.t0_day, n_days, sub.forcing, sub.site_of_member, sub.scal,
                                sub.lay, sub.u0)
    finally:
        emul.use_wide(False)
    assert np.array_equal(st, st_o)
    assert np.array_equal(cnt, cnt_o), "accepted / rejected / RHS-evaluation counters differ"
    same = (got == ref) | (np.isnan(got) & np.isnan(ref))
    assert same.all(), f"{(~same).any(axis=(1, 2)).sum()} of {sub.M} members differ from the oracle"
    return ref, cnt


def test_lane_kernel_dav2020_isotopes(dav_full):
    """N = 5 (6 members per warp, 2 surplus lanes), isotopes, partial last warp."""
    check(en.make_ensemble(dav_full, 16), 8, emul.LANES)


def test_lane_kernel_stays_on_the_main_paths_of_exp_and_log(dav_full):
    """No lane may compute with NaN garbage: results would not show it (such lanes never commit), but every exp / log of
    the warp would go through the out-of-line cold paths (round 2: the surplus lanes of a warp once did, after the
    callback core moved to one lane per member; 2x the run time on the GPU).  Healthy DAV2020 members take the cold path
    only for saturated or dry layers: far less than once per RHS evaluation and lane."""
    import ctypes as C
    e = en.make_ensemble(dav_full, 12)
    L = emul.lib()
    L.lwfb_emul_cold_calls.restype = C.c_long
    L.lwfb_emul_cold_calls()
    _, st, cnt = emul.run(emul.LANES, e.N, e.iso, e.t0_day, 6, e.forcing, e.site_of_member, e.scal, e.lay, e.u0)
    cold = L.lwfb_emul_cold_calls()
    assert (st == 0).all()
    assert cold < 0.01 * cnt[:, 2].sum() * e.N, f"{cold} cold-path calls for {cnt[:, 2].sum()} RHS evaluations"


def test_lane_kernel_bea_no_isotopes(bea):
    """N = 7 (4 members per warp, 4 surplus lanes), no isotopes."""
    check(en.make_ensemble(bea, 6), 10, emul.LANES)


def test_lane_kernel_isotopes_seven_layers(isobea):
    check(en.make_ensemble(isobea, 5), 6, emul.LANES)


def infexp(lw, n_layers):
    m = lw.loadSPAC(inputs("infiltrationSaturationINFEXP1"), "infiltrationSaturationINFEXP1", simulate_isotopes=False,
                    Δz_thickness_m=[2.1 / n_layers] * n_layers,
                    IC_soil=dict(PSIM_init_kPa=-6.3, delta18O_init_permil=-13.0, delta2H_init_permil=-95.0),
                    root_distribution=dict(beta=0.90, z_rootMax_m=None))
    return lw.setup(m, requested_tspan=(0, 300))


def test_lane_kernel_saturation_bypass_groundwater(lw):
    """INFEXP1 on a 21-layer grid during the infiltration pulse: saturated layers (log 0 / pow 0 paths), the
    bottom-up INFLOW limiter, bypass flow (BYPAR = 1), groundwater feedback (GSC = 0.01) and hundreds of
    rejected steps."""
    s = infexp(lw, 21)
    assert s.N == 21
    e = en.batched_setup(s, params={"DRAIN": np.array([1.0, 0.6, 0.3, 0.8]), "BYPAR": np.array([0., 1., 1., 0.]),
                                    "QFPAR": np.array([0.3, 0.3, 0.1, 0.3]), "QFFC": np.array([0.2, 0.2, 0.4, 0.2])})
    assert (e.scal[:, lo.S.GSC] > 0).all()
    e.t0_day, e.n_days = 226, 12
    ref, cnt = check(e, e.n_days, emul.LANES)
    ix = lo.StateIndex(e.N, e.iso)
    W = ref[:, :, ix.SWATI_mm] / e.lay[:, None, lo.L.SWATMAX, :]
    assert (W >= 1.0).any(), "the window must contain saturated layers"
    assert (cnt[:, 1] > 20).all(), "the window must contain rejected steps"
    # the same window in work quanta of 11 passes with regrouping: checkpoints fall between rejected steps and their
    # retries, inside the infiltration pulse, with groundwater feedback (GWAT partial sums) and 21 lanes per member
    got, st, cnt_q = emul.run(emul.LANES, e.N, e.iso, e.t0_day, e.n_days, e.forcing, e.site_of_member, e.scal, e.lay, e.u0,
                              seg_len=11, resort=True)
    assert (st == 0).all() and np.array_equal(cnt_q, cnt) and np.array_equal(got, ref, equal_nan=True)
    # the general instantiation in multi-warp blocks: the block-level callback phase serves members of other warps
    # (thread k of the block = member k), the stage barriers and votes span the block
    got, st, cnt_b = emul.run(emul.LANES, e.N, e.iso, e.t0_day, e.n_days, e.forcing, e.site_of_member, e.scal, e.lay, e.u0,
                              warps_per_block=4)
    assert (st == 0).all() and np.array_equal(cnt_b, cnt) and np.array_equal(got, ref, equal_nan=True)


@pytest.mark.parametrize("N", [3, 4, 8, 16, 32])
def test_lane_kernel_group_geometries(lw, N):
    """32/N members per warp, 1..3 accumulator slots per lane, N = 32 fills the warp."""
    m = lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True, Δz_thickness_m=aligned_grid(N),
                    root_distribution=dict(beta=0.97, z_rootMax_m=None),
                    IC_soil=dict(PSIM_init_kPa=-6.0, delta18O_init_permil=-12.0, delta2H_init_permil=-90.0))
    s = lw.setup(m, ε=0.001 if N > 16 else 0.01, requested_tspan=(100, 104))
    assert s.N == N
    e = en.make_ensemble(s, 5)
    ref, cnt = check(e, 4, emul.LANES)
    # and in work quanta (the checkpoint holds 1..3 accumulator derivatives per lane, depending on the geometry)
    got, st, cnt_q = emul.run(emul.LANES, e.N, e.iso, e.t0_day, 4, e.forcing, e.site_of_member, e.scal, e.lay, e.u0,
                              seg_len=7, resort=True)
    assert (st == 0).all() and np.array_equal(cnt_q, cnt) and np.array_equal(got, ref, equal_nan=True)


def test_member_kernel_generic_layer_count(lw, dav_full):
    """Thread-per-member fallback kernel: N = 42 (outside the lane kernel's range) and N = 5."""
    e = en.batched_setup(infexp(lw, 42), params={"DRAIN": np.array([1.0, 0.3])})
    check(e, 20, emul.MEMBER)
    check(en.make_ensemble(dav_full, 6), 10, emul.MEMBER)


def test_wide_kernel_saturation_bypass_groundwater(lw):
    """Warp-per-member kernel (N > 32, lwfb90_wide.cuh): INFEXP1 on its 42-layer grid through the infiltration pulse -
    saturated layers, the bottom-up INFLOW limiter (lane 0 walks the chain only when a warp vote says a layer limits),
    bypass flow, groundwater feedback, ITER's clamped chain, hundreds of rejected steps."""
    s = infexp(lw, 42)
    assert s.N == 42
    e = en.batched_setup(s, params={"DRAIN": np.array([1.0, 0.6, 0.3, 0.8]), "BYPAR": np.array([0., 1., 1., 0.]),
                                    "QFPAR": np.array([0.3, 0.3, 0.1, 0.3]), "QFFC": np.array([0.2, 0.2, 0.4, 0.2])})
    e.t0_day, e.n_days = 226, 6
    ref, cnt = check(e, e.n_days, emul.WIDE)
    ix = lo.StateIndex(e.N, e.iso)
    assert (ref[:, :, ix.SWATI_mm] / e.lay[:, None, lo.L.SWATMAX, :] >= 1.0).any(), "the window must contain saturated layers"
    assert (cnt[:, 1] > 10).all(), "the window must contain rejected steps"


@pytest.mark.parametrize("N", [40, 70])
def test_wide_kernel_isotopes_deep_profiles(lw, N):
    """Warp-per-member kernel with isotopes: 40 layers (the layer count of golden G4) and 70 (three layers per lane for
    some lanes, two for others), scores accumulated at observation days, du_NTFLI left behind, one aborting member."""
    m = lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True, Δz_thickness_m=aligned_grid(N),
                    root_distribution=dict(beta=0.97, z_rootMax_m=None),
                    IC_soil=dict(PSIM_init_kPa=-6.0, delta18O_init_permil=-12.0, delta2H_init_permil=-90.0))
    s = lw.setup(m, ε=0.001, requested_tspan=(100, 103))
    assert s.N == N
    e = en.make_ensemble(s, 3)
    e.u0[2, 7] = np.nan                       # member 2: NaN in SWATI -> Unstable at its first step
    nd = 3 if N == 40 else 1
    orc.use("elem")
    obs_day, obs_layer = np.array([0, 1]), np.array([0, N // 2, N - 1])
    rng = np.random.default_rng(2)
    th, d18 = rng.uniform(0.1, 0.4, (2, 3)), rng.uniform(-14, -8, (2, 3))
    ref, st_o, cnt_o, sums_o, nt_o = [], [], [], [], []
    for k in range(e.M):
        r = orc.run_member(e.N, e.iso, e.t0_day, nd, e.forcing[0], e.scal[k], e.lay[k], e.u0[k])
        ref.append(r[0]); st_o.append(r[1]); cnt_o.append(r[2]); nt_o.append(r[3])
    emul.use_wide(True)
    try:
        emul.set_observations(nd, obs_day, obs_layer, th, d18)
        got, st, cnt = emul.run(emul.WIDE, e.N, e.iso, e.t0_day, nd, e.forcing, e.site_of_member, e.scal, e.lay, e.u0)
        sums = emul.score_sums(e.M)
        nt = emul.get_ntfli(e.M, e.N)
        emul.set_observations(nd)
    finally:
        emul.use_wide(False)
    assert list(st) == st_o and st_o[2] != 0 and st_o[0] == 0
    assert np.array_equal(cnt, np.array(cnt_o)) and np.array_equal(got, np.array(ref), equal_nan=True)
    assert np.array_equal(nt[:2], np.array(nt_o)[:2], equal_nan=True)
    # the sums: the same additions in the same order as a scorer walking the record (days outer, layers inner)
    ix = lo.StateIndex(N, True)
    for k in range(2):
        a = np.zeros(7); b = np.zeros(7)
        for j, dday in enumerate(obs_day):
            for l, il in enumerate(obs_layer):
                sw = got[k, dday, ix.SWATI_mm][il]
                ths, thr, smax = e.lay[k, lo.L.THSAT, il], e.lay[k, lo.L.THR, il], e.lay[k, lo.L.SWATMAX, il]
                w = min((ths * sw / smax - thr) / (ths - thr), 1.0)
                for acc, sim_, o in ((a, w * (ths - thr) + thr, th[j, l]), (b, got[k, dday, ix.SWATI_d18O][il], d18[j, l])):
                    err = sim_ - o
                    acc += np.array([1, err * err, o, sim_, o * o, sim_ * sim_, o * sim_])
        assert np.allclose(sums[k, 0], a, rtol=1e-14) and np.allclose(sums[k, 1], b, rtol=1e-14)


def test_wide_kernel_rhs_probe(lw):
    """One evaluation of f(u, t) by the warp-per-member kernel's own RHS (the probe behind lwfb_eval_rhs for N > 32)."""
    e = en.batched_setup(infexp(lw, 42), params={"DRAIN": np.linspace(0.1, 1, 6), "BYPAR": np.array([0., 1., 0., 1., 0., 1.])})
    N, M = e.N, e.M
    rng = np.random.default_rng(5)
    ix = lo.StateIndex(N, e.iso)
    smax = e.lay[:, lo.L.SWATMAX]
    swati = e.u0[:, ix.SWATI_mm] * rng.uniform(0.5, 1.3, (M, N))
    swati[0] = smax[0] * 1.01; swati[3, 0] = -0.1
    gwat = rng.uniform(0, 5, M); gwat[1] = 0.0
    frozen = np.concatenate([rng.uniform(0, 30, (M, 1)), rng.uniform(0, 2, (M, 1)), rng.uniform(0, 1, (M, 1)),
                             rng.uniform(-0.01, 0.5, (M, N))], axis=1)
    ntfli = rng.normal(0, 20, (M, N)); ntfli[2] = np.nan
    orc.use("elem")
    emul.use_wide(True)
    try:
        du = emul.eval_rhs(N, e.iso, e.forcing, e.site_of_member, e.scal, e.lay, e.u0, 0.3712, swati, gwat, frozen, ntfli,
                           kernel=emul.WIDE)
    finally:
        emul.use_wide(False)
    for m in range(M):
        ref, _ = orc.eval_rhs(N, e.scal[m], e.lay[m], 0.3712, swati[m], gwat[m], frozen[m], ntfli[m])
        assert np.array_equal(du[m], ref, equal_nan=True), m


def test_lane_kernel_placement_does_not_change_results(dav_full):
    """Members may be placed in any slot (lwfb_set_placement): a random permutation gives the same records."""
    e = en.make_ensemble(dav_full, 15)
    orc.use("elem")
    ref, st_o, cnt_o = orc.run(e.N, e.iso, e.t0_day, 6, e.forcing, e.site_of_member, e.scal, e.lay, e.u0, n_threads=8)
    perm = np.random.default_rng(3).permutation(e.M).astype(np.int32)
    got, st, cnt = emul.run(emul.LANES, e.N, e.iso, e.t0_day, 6, e.forcing, e.site_of_member, e.scal, e.lay, e.u0, perm=perm)
    assert np.array_equal(st, st_o) and np.array_equal(cnt, cnt_o)
    assert np.array_equal(got, ref, equal_nan=True)
    # empty slots (placement entry -1): members spread thinly over more warps than they need
    holes = np.full(3 * e.M, -1, dtype=np.int32)
    holes[np.sort(np.random.default_rng(4).choice(3 * e.M, e.M, replace=False))] = perm
    got, st, cnt = emul.run(emul.LANES, e.N, e.iso, e.t0_day, 6, e.forcing, e.site_of_member, e.scal, e.lay, e.u0, perm=holes)
    assert np.array_equal(st, st_o) and np.array_equal(cnt, cnt_o)
    assert np.array_equal(got, ref, equal_nan=True)


def test_lane_kernel_work_quanta_with_regrouping(dav_full, isobea):
    """A launch ends after a fixed number of passes of the block loop; every member checkpoints where it is (also in
    the middle of a day or between a rejected step and its retry), the unfinished members are regrouped by progress
    and resumed (what lwfb_run does).  Must equal the oracle's single pass for any quantum, down to one pass."""
    for sim, M, nd, quantum, resort in ((dav_full, 14, 11, 37, True), (dav_full, 7, 4, 1, True), (isobea, 5, 7, 16, True),
                                        (dav_full, 13, 6, 5, False)):
        e = en.make_ensemble(sim, M)
        orc.use("elem")
        ref, st_o, cnt_o = orc.run(e.N, e.iso, e.t0_day, nd, e.forcing, e.site_of_member, e.scal, e.lay, e.u0, n_threads=8)
        got, st, cnt = emul.run(emul.LANES, e.N, e.iso, e.t0_day, nd, e.forcing, e.site_of_member, e.scal, e.lay, e.u0,
                                seg_len=quantum, resort=resort)
        assert np.array_equal(st, st_o) and np.array_equal(cnt, cnt_o)
        assert np.array_equal(got, ref, equal_nan=True)


def test_lane_kernel_aborting_members(dav_full):
    """Members that stop early (retcode MaxIters after 120 attempted steps, Unstable from a NaN state) keep the records
    they reached and report NaN afterwards, in a single launch and across work quanta (the abort happens in one launch,
    later launches must leave the member alone)."""
    e = en.make_ensemble(dav_full, 13)
    e.u0[4, lo.StateIndex(e.N, e.iso).SWATI_mm.start + 2] = np.nan
    orc.use("elem")
    kw = dict(maxiters=120)
    ref, st_o, cnt_o = orc.run(e.N, e.iso, e.t0_day, 9, e.forcing, e.site_of_member, e.scal, e.lay, e.u0, n_threads=8, **kw)
    assert (st_o == 3).sum() >= 10 and st_o[4] == 1, st_o            # LWFB_ST_MAXITERS / LWFB_ST_UNSTABLE
    assert np.isnan(ref[:, -1]).all() and not np.isnan(ref[0, 2]).any()
    for quantum, resort in ((0, False), (23, True), (1, True)):
        got, st, cnt = emul.run(emul.LANES, e.N, e.iso, e.t0_day, 9, e.forcing, e.site_of_member, e.scal, e.lay, e.u0,
                                seg_len=quantum, resort=resort, **kw)
        assert np.array_equal(st, st_o) and np.array_equal(cnt, cnt_o)
        assert np.array_equal(got, ref, equal_nan=True)
    got, st, cnt = emul.run(emul.MEMBER, e.N, e.iso, e.t0_day, 9, e.forcing, e.site_of_member, e.scal, e.lay, e.u0, **kw)
    assert np.array_equal(st, st_o) and np.array_equal(cnt, cnt_o) and np.array_equal(got, ref, equal_nan=True)


def test_lane_kernel_full_thread_blocks(dav_full, bea):
    """What the GPU runs: thread blocks of 8 warps (and 2, 4) in lock step — block-uniform votes (__syncthreads_or), a
    barrier per Runge-Kutta stage, warps without a stepping member idling through the barriers of a pass, bounded
    waiting at the day boundary across the warps of a block, the last block partly empty.  The emulator matches block
    barriers as the hardware does and aborts on a mismatch (a hang on the GPU).  One launch and work quanta."""
    for sim, M, nd, wpb, quantum, resort in ((dav_full, 40, 6, 8, 0, False), (dav_full, 70, 5, 8, 19, True),
                                             (bea, 21, 6, 2, 5, True), (dav_full, 100, 4, 4, 1, True)):
        e = en.make_ensemble(sim, M)
        orc.use("elem")
        ref, st_o, cnt_o = orc.run(e.N, e.iso, e.t0_day, nd, e.forcing, e.site_of_member, e.scal, e.lay, e.u0, n_threads=8)
        got, st, cnt = emul.run(emul.LANES, e.N, e.iso, e.t0_day, nd, e.forcing, e.site_of_member, e.scal, e.lay, e.u0,
                                seg_len=quantum, resort=resort, warps_per_block=wpb)
        assert np.array_equal(st, st_o) and np.array_equal(cnt, cnt_o)
        assert np.array_equal(got, ref, equal_nan=True)


def test_lane_kernel_stop_flag_ends_launches_early(dav_full):
    """The launch protocol of lwfb_run as the __global__ wrapper drives it: blocks run concurrently (host threads), the
    first block that uses up its quantum when no block is left to take its slot raises the stop flag, every block
    checkpoints at its next pass (threads of a block may see the flag one pass apart - the block vote decides), the
    front block has no budget of its own.  Where a launch is cut depends on thread timing; the records must not."""
    for M, nd, wpb, quantum, threads in ((120, 5, 2, 9, 4), (200, 4, 1, 7, 3), (90, 4, 8, 13, 2)):
        e = en.make_ensemble(dav_full, M)
        orc.use("elem")
        ref, st_o, cnt_o = orc.run(e.N, e.iso, e.t0_day, nd, e.forcing, e.site_of_member, e.scal, e.lay, e.u0, n_threads=8)
        got, st, cnt = emul.run(emul.LANES, e.N, e.iso, e.t0_day, nd, e.forcing, e.site_of_member, e.scal, e.lay, e.u0,
                                seg_len=quantum, resort=2, warps_per_block=wpb, n_threads=threads)
        assert np.array_equal(st, st_o) and np.array_equal(cnt, cnt_o)
        assert np.array_equal(got, ref, equal_nan=True)


def test_emulator_rejects_divergent_warp_primitives():
    """The emulator aborts when lanes meet at different warp primitives; a well-formed run reports none.  (The
    abort path itself is exercised by the build of every kernel change: a divergent shuffle never gets here.)"""
    assert os.path.exists(emul.build())


def _score_sums_numpy(rec, lay, iso, N, obs_day, obs_layer, obs_theta, obs_d18o):
    """Sequential restatement of score_add over (days outer, layers inner) on a daily record [M][n_days+1][L]."""
    M = rec.shape[0]
    out = np.zeros((M, 2, 7))
    for m in range(M):
        for k, dday in enumerate(obs_day):
            for l, i in enumerate(obs_layer):
                ths, thr, smax = lay[m, lo.L.THSAT, i], lay[m, lo.L.THR, i], lay[m, lo.L.SWATMAX, i]
                dth = ths - thr
                x = (ths * rec[m, dday, i] / smax - thr) / dth
                w = x if not (x >= 1.0) else 1.0
                sims = [(w * dth + thr, obs_theta[k, l])]
                if iso:
                    sims.append((rec[m, dday, N + i], obs_d18o[k, l]))
                for v, (sim, o) in enumerate(sims):
                    if o != o:
                        continue
                    e = sim - o
                    a = out[m, v]
                    a[0] += 1; a[1] += e * e; a[2] += o; a[3] += sim; a[4] += o * o; a[5] += sim * sim; a[6] += o * sim
    return out


@pytest.mark.parametrize("seg_len", [0, 7])
def test_score_sums_accumulated_by_the_integrator(dav_full, seg_len):
    """lwfb_set_observations: the sums the lane kernel accumulates at the observation days (no record needed) equal,
    bit for bit, the sums formed from the daily record afterwards - in one launch and in work quanta, for members
    that finish and for one that aborts (its unreached observation days count as NaN, like its NaN records)."""
    ens = en.make_ensemble(dav_full, 9)
    nd = 12
    ens.u0[4, 2] = np.nan            # member 4: NaN in SWATI -> Unstable at its first step
    obs_day = np.array([0, 3, 4, 9, 12], dtype=np.int32)
    obs_layer = np.array([0, 2, 4], dtype=np.int32)
    rng = np.random.default_rng(5)
    th = 0.2 + 0.05 * rng.random((5, 3)); d18 = -11 + rng.random((5, 3))
    th[1, 2] = np.nan; d18[3, 0] = np.nan
    emul.set_observations(nd, obs_day, obs_layer, th, d18)
    try:
        rec, st, cnt = emul.run(emul.LANES, ens.N, ens.iso, ens.t0_day, nd, ens.forcing, ens.site_of_member, ens.scal,
                                ens.lay, ens.u0, seg_len=seg_len, resort=2 if seg_len else False, warps_per_block=2)
        got = emul.score_sums(ens.M)
    finally:
        emul.set_observations(nd)
    assert st[4] != 0 and (np.delete(st, 4) == 0).all()
    want = _score_sums_numpy(rec, ens.lay, ens.iso, ens.N, obs_day, obs_layer, th, d18)
    assert np.array_equal(got, want, equal_nan=True)
    assert np.isnan(got[4, :, 1]).all() and np.isfinite(np.delete(got, 4, axis=0)).all()
    # the thread-per-member kernel accumulates the same sums
    emul.set_observations(nd, obs_day, obs_layer, th, d18)
    try:
        rec2, _, _ = emul.run(emul.MEMBER, ens.N, ens.iso, ens.t0_day, nd, ens.forcing, ens.site_of_member, ens.scal,
                              ens.lay, ens.u0)
        got2 = emul.score_sums(ens.M)
    finally:
        emul.set_observations(nd)
    assert np.array_equal(rec2, rec, equal_nan=True) and np.array_equal(got2, want, equal_nan=True)


@pytest.mark.parametrize("which", ["dav_full", "bea"])
def test_lane_kernel_rhs_probe(which, request):
    """Unit test of the LANE kernel's right-hand side (the probe behind lwfb_eval_rhs): one evaluation of f(u, t) by
    the production kernel's own stage code against the oracle, for perturbed, clamped and saturated states and a NaN
    du_NTFLI cache."""
    sim = request.getfixturevalue(which)
    ens = en.make_ensemble(sim, 14)
    N, M = ens.N, ens.M
    rng = np.random.default_rng(5)
    ix = lo.StateIndex(N, ens.iso)
    smax = ens.lay[:, lo.L.SWATMAX]
    swati = ens.u0[:, ix.SWATI_mm] * rng.uniform(0.5, 1.3, (M, N))
    swati[::7] = smax[::7] * 1.01
    swati[3::11, 0] = -0.1
    gwat = rng.uniform(0, 5, M); gwat[::5] = 0.0
    frozen = np.concatenate([rng.uniform(0, 30, (M, 1)), rng.uniform(0, 2, (M, 1)), rng.uniform(0, 1, (M, 1)),
                             rng.uniform(-0.01, 0.5, (M, N))], axis=1)
    ntfli = rng.normal(0, 20, (M, N)); ntfli[::9] = np.nan
    orc.use("elem")
    for t in (0.0, 0.3712, 10.999999):
        du = emul.eval_rhs(N, ens.iso, ens.forcing, ens.site_of_member, ens.scal, ens.lay, ens.u0, t, swati, gwat, frozen, ntfli,
                           warps_per_block=2)
        for m in range(M):
            ref, _ = orc.eval_rhs(N, ens.scal[m], ens.lay[m], t, swati[m], gwat[m], frozen[m], ntfli[m])
            assert np.array_equal(du[m], ref, equal_nan=True), (t, m, du[m], ref)


def test_warm_ntfli_cache_second_run(dav_full):
    """a second simulate! on the same problem starts from the du_NTFLI the first run left behind.  The
    lane kernel with lwfb_set_ntfli0 reproduces the oracle's warm-cache run bit for bit.  (Whether the records differ from
    a cold run depends on whether ITER's limit binds in the first step; the RHS probe test covers the NaN / finite cases.)"""
    ens = en.make_ensemble(dav_full, 6)
    nd = 5
    orc.use("elem")
    cold = [orc.run_member(ens.N, ens.iso, 0, nd, ens.forcing[0], ens.scal[m], ens.lay[m], ens.u0[m]) for m in range(ens.M)]
    warm = [orc.run_member(ens.N, ens.iso, 0, nd, ens.forcing[0], ens.scal[m], ens.lay[m], ens.u0[m], ntfli_init=cold[m][3])
            for m in range(ens.M)]
    rec1, st1, cnt1 = emul.run(emul.LANES, ens.N, ens.iso, 0, nd, ens.forcing, ens.site_of_member, ens.scal, ens.lay, ens.u0)
    nt1 = emul.get_ntfli(ens.M, ens.N)
    assert np.array_equal(nt1, np.stack([c[3] for c in cold]))
    emul.set_ntfli0(nt1)
    try:
        rec2, st2, cnt2 = emul.run(emul.LANES, ens.N, ens.iso, 0, nd, ens.forcing, ens.site_of_member, ens.scal, ens.lay, ens.u0)
    finally:
        emul.set_ntfli0(None)
    assert np.array_equal(rec2, np.stack([w[0] for w in warm]), equal_nan=True)
    assert np.array_equal(cnt2, np.stack([w[2] for w in warm]))
    assert np.array_equal(emul.get_ntfli(ens.M, ens.N), np.stack([w[3] for w in warm]))
