// lwfb90_lanes.cuh — the production integrator: one LANE per soil layer.
//
// Mapping: a member with N layers (3 <= N <= 32) occupies N consecutive lanes of a warp ("group");
// a warp holds floor(32/N) members, the remaining lanes idle.  Lane i of a group owns layer i:
// its 14 hydraulic constants, its SWATI component with the Runge-Kutta bookkeeping, its delta18O /
// delta2H, and up to EMAX of the 9 member-level ODE components (GWAT and the 8 integrated flow
// accumulators are dealt round-robin to the lanes of the group).  Everything lives in registers; there
// is no shared memory and no per-thread stack in the stepping loop.
//
// Why this mapping (see profiles/r01_*): with one member per THREAD the layer loops must be unrolled
// to keep per-layer arrays in registers, the inlined exp/log/pow bodies then make one RHS ~9000
// instructions (~140 KB of SASS), far beyond the instruction cache, and with 255 registers only 8 warps
// per SM are resident: ncu showed the warps stalled on instruction fetch ~10 cycles per issued
// instruction and the FP64 pipe 6 % busy.  One lane per layer makes the code naturally rolled (the
// per-layer body exists once), needs ~1/5 of the registers, and turns cross-layer couplings into warp
// shuffles:
//   * Darcy flux needs psi_t and log K of the layer below          -> __shfl_down
//   * ITER's min-chain, the bypass/downslope sums, the error norm   -> N-step broadcast chains, evaluated
//     in exactly the oracle's order so that results stay bit-identical
//   * INFLOW's bottom-up oversaturation limiter                     -> skipped by a warp vote when no layer
//     limits (the common case), otherwise an N-step chain
// Member-level scalars (t, dt, controller state, frozen daily fluxes) are replicated in the lanes of a
// group; they evolve identically because every input is a broadcast value.  The daily callback runs
// warp-synchronously (lanes wait at the day boundary for the slowest member of the warp).
#pragma once
#include "lwfb90_device.cuh"

namespace lwfb {

// CQ[ks][p] = a(ks+1+p, ks): contribution of k_ks to the input of stage ks+1+p (stage 7 = the new state)
__device__ constexpr double lCQ[7][5] = {
    {0, 0, 0, 0, 0},
    {-0.008480655492356989, 2.8971530571054935, 5.325864828439257, 5.86145544294642, 0.09646076681806523},   // k1 -> stages 3..7
    {0.335480655492357, -6.359448489975075, -11.748883564062828, -12.92096931784711, 0.01},                  // k2 -> stages 3..7
    {4.3622954328695815, 7.4955393428898365, 8.159367898576159, 0.4798896504144996, 0},                       // k3 -> stages 4..7
    {-0.09249506636175525, -0.071584973281401, 1.379008574103742, 0, 0},                                      // k4 -> stages 5..7
    {-0.028269050394068383, -3.290069515436081, 0, 0, 0},                                                     // k5 -> stages 6..7
    {2.324710524099774, 0, 0, 0, 0}};                                                                         // k6 -> stage 7
__device__ constexpr double lBT[8] = {0, -0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
                                      -0.1447110071732629, 0.5823571654525552, -0.45808210592918697, 0.015151515151515152};
__device__ constexpr double lC[7] = {0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0};   // time of the stage evaluated at s
__device__ constexpr int lAccPos[8] = {LWFB_A_FLOW, LWFB_A_SEEP, LWFB_A_SRFL, LWFB_A_SLFL,
                                       LWFB_A_BYFL, LWFB_A_DSFL, LWFB_A_GWFL, LWFB_A_VRFLN};

#define LFULL 0xffffffffu

template <int NMAX, int EMAX>
__device__ void lanes_body(const DevData& d, const int lane, const int64_t warp_g) {
  constexpr int NC = 1 + EMAX;                       // ODE components per lane
  const int N = d.N;
  const int mpw = 32 / N;                            // members per warp
  const int64_t M = d.M;
  int g = lane / N;
  int i = lane - g * N;
  const bool lane_valid = g < mpw;
  if (!lane_valid) { g = 0; i = 0; }                 // surplus lanes shadow lane 0 and never commit
  const int base = g * N;
  const int64_t m_raw = warp_g * mpw + g;
  const bool live = lane_valid && m_raw < M;
  const int64_t m = (m_raw < M) ? m_raw : M - 1;
  const bool top = (i == 0), bot = (i == N - 1);
  const int L = d.L;
  const bool iso = d.iso != 0;
  const SIdx ix(N, d.iso);
  const ParamView pv{N, m, M, d.scal, d.lay};
  const double* __restrict__ forc = d.forc + (int64_t)__ldg(d.site + m) * d.n_rows * LWFB_F;
  auto U0 = [&](int j) { return __ldg(d.u0 + (int64_t)j * M + m); };
  auto bcast = [&](double v, int layer) { return __shfl_sync(LFULL, v, base + layer); };

  // ---- per-lane constants (own layer i, and what the flux to the layer below needs)
  const double THSAT = pv.Lp(LWFB_L_THSAT, i), SWATMAX = pv.Lp(LWFB_L_SWATMAX, i), THR = pv.Lp(LWFB_L_THR, i);
  const double NPAR = pv.Lp(LWFB_L_NPAR, i), MPAR = pv.Lp(LWFB_L_MPAR, i);
  const double E1 = -1.0 / MPAR, E2 = 1.0 / NPAR, E3 = NPAR / (NPAR - 1.0);
  const double TORT = pv.Lp(LWFB_L_TORT, i), CPSI = 9.81 * (-1.0 / pv.Lp(LWFB_L_ALPHA, i));
  const double KSAT = pv.Lp(LWFB_L_KSAT, i), PSIG = pv.Lp(LWFB_L_PSIG, i), THICK = pv.Lp(LWFB_L_THICK, i);
  const double STONEF = pv.Lp(LWFB_L_STONEF, i), INFRAC = pv.Lp(LWFB_L_INFRAC, i);
  const int idn = bot ? i : i + 1;
  const double THICK_dn = pv.Lp(LWFB_L_THICK, idn), KSAT_dn = pv.Lp(LWFB_L_KSAT, idn), STONEF_dn = pv.Lp(LWFB_L_STONEF, idn);
  const double DSWMAX = pv.S(LWFB_S_DSWMAX);
  const double CITER = 0.01 * DSWMAX * THICK * THSAT * (1 - STONEF);
  const double DPSIMAX = pv.S(LWFB_S_DPSIMAX), DTIMAX = pv.S(LWFB_S_DTIMAX), DRAIN = pv.S(LWFB_S_DRAIN),
               GSC = pv.S(LWFB_S_GSC), GSP = pv.S(LWFB_S_GSP), IMPERV = pv.S(LWFB_S_IMPERV);
  const bool BYPAR = pv.S(LWFB_S_BYPAR) == 1.0;
  const int QLAYER = (int)pv.S(LWFB_S_QLAYER);
  const bool SLOPE = !(pv.S(LWFB_S_LENGTH_SLOPE) == 0 || pv.S(LWFB_S_DSLOPE) == 0);
  const double WETF = BYPAR ? pv.Lp(LWFB_L_WETF, i) : 0.0;
  const double DZ_up = top ? 1.0 : pv.Lp(LWFB_L_DZ, i - 1), DZ_dn = bot ? 1.0 : pv.Lp(LWFB_L_DZ, i);

  // ---- ODE components of this lane: c = 0 own SWATI; c >= 1: member-level component e = i + (c-1)*N
  //      (e = 0 GWAT, 1..8 the integrated accumulators flow, seep, srfl, slfl, byfl, dsfl, gwfl, vrfln)
  double y[NC], yprev[NC], k1[NC], kk_[NC];
  bool cvalid[NC];
  int cext[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) {
    cext[c] = (c == 0) ? -1 : i + (c - 1) * N;
    cvalid[c] = (c == 0) || (cext[c] < 9);
    k1[c] = 0; kk_[c] = 0; yprev[c] = 0;
  }
  DayState ds;
  for (int a = 0; a < LWFB_NACCUM; a++) ds.acc[a] = U0(ix.ACC + a);
  y[0] = U0(ix.SWATI + i);
#pragma unroll
  for (int c = 1; c < NC; c++) {
    const int e = cext[c];
    y[c] = !cvalid[c] ? 0.0 : (e == 0 ? U0(ix.GWAT) : ds.acc[lAccPos[(e - 1) & 7]]);
  }
  ds.INTS = U0(ix.INTS); ds.INTR = U0(ix.INTR); ds.SNOW = U0(ix.SNOW); ds.CC = U0(ix.CC); ds.SNOWLQ = U0(ix.SNOWLQ);
  double TRANIst = U0(ix.TRANI + i), RWUst = U0(ix.RWU), XYmm = U0(ix.XYLEM);
  double dS[2] = {0, 0}, dG[2] = {0, 0}, dRWU[2] = {0, 0}, dXY[2] = {0, 0}, dISN[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  if (iso) {
#pragma unroll
    for (int s = 0; s < 2; s++) {
      dS[s] = U0(ix.SWATI + (s + 1) * N + i);
      dG[s] = U0(ix.GWAT + 1 + s); dRWU[s] = U0(ix.RWU + 1 + s); dXY[s] = U0(ix.XYLEM + 1 + s);
      dISN[0][s] = U0(ix.INTS + 1 + s); dISN[1][s] = U0(ix.INTR + 1 + s); dISN[2][s] = U0(ix.SNOW + 1 + s);
    }
  }
  double TRANI = NAN, ntfli = NAN;          // frozen uptake of the layer; du_NTFLI of the previous RHS call
  double RNET = NAN, SMLT = NAN, SLVP = NAN, TA = NAN;

  // value of member-level component e (held by lane e % N in slot 1 + e / N), broadcast to the group
  auto ext_bcast = [&](const double* v, int e) {
    const int slot = 1 + e / N;
    double mine = v[1];
#pragma unroll
    for (int c = 2; c < NC; c++) mine = (slot == c) ? v[c] : mine;
    return bcast(mine, e % N);
  };

  // full state record in the reference's order; lane i writes the entries of layer i, lane 0 the scalars
  auto write_record = [&](double* __restrict__ dst) {
    if (!live) return;
    auto W = [&](int j, double v) { dst[(int64_t)j * M + m] = v; };
    W(ix.SWATI + i, y[0]);
    W(ix.TRANI + i, TRANIst);
    if (iso) {
#pragma unroll
      for (int s = 0; s < 2; s++) {
        W(ix.SWATI + (s + 1) * N + i, dS[s]);
        W(ix.TRANI + (s + 1) * N + i, U0(ix.TRANI + (s + 1) * N + i));
      }
    }
#pragma unroll
    for (int q = 0; q < 3; q++) W(ix.AUX + q * N + i, U0(ix.AUX + q * N + i));
#pragma unroll
    for (int c = 1; c < NC; c++)
      if (cvalid[c]) W(cext[c] == 0 ? ix.GWAT : ix.ACC + lAccPos[(cext[c] - 1) & 7], y[c]);
    if (top) {
      W(ix.INTS, ds.INTS); W(ix.INTR, ds.INTR); W(ix.SNOW, ds.SNOW);
      W(ix.RWU, RWUst); W(ix.XYLEM, XYmm); W(ix.CC, ds.CC); W(ix.SNOWLQ, ds.SNOWLQ);
      if (iso) {
#pragma unroll
        for (int s = 0; s < 2; s++) {
          W(ix.GWAT + 1 + s, dG[s]); W(ix.RWU + 1 + s, dRWU[s]); W(ix.XYLEM + 1 + s, dXY[s]);
          W(ix.INTS + 1 + s, dISN[0][s]); W(ix.INTR + 1 + s, dISN[1][s]); W(ix.SNOW + 1 + s, dISN[2][s]);
        }
      }
      for (int a = 0; a < LWFB_NACCUM; a++) {
        bool integrated = false;
#pragma unroll
        for (int q = 0; q < 8; q++) integrated |= (a == lAccPos[q]);
        if (!integrated) W(ix.ACC + a, ds.acc[a]);
      }
    }
  };
  auto prec_delta = [&](double t, double* dp) {
    int r = (int)ceil(t - 0.5);
    r = r < 0 ? 0 : (r > d.n_rows - 1 ? d.n_rows - 1 : r);
    dp[0] = __ldg(forc + (int64_t)r * LWFB_F + LWFB_F_D18O);
    dp[1] = __ldg(forc + (int64_t)r * LWFB_F + LWFB_F_D2H);
  };

  const double qmin = 0.2, qmax = 10.0, gamma = 0.9, beta1 = 0.14, beta2 = 0.08, qoldinit = 1e-4;
  const double t0 = (double)d.t0_day, tend = (double)(d.t0_day + d.n_days);
  double t = t0, tprev = t0, dt = d.dt_init, dtpropose = d.dt_init, qold = qoldinit, q11 = 1.0;
  const double dtmin_opt = fmax(eps_of(tend), eps_of(t0));
  long long naccept = 0, nreject = 0, iter = 0, nf = 0;
  bool accept = false, reeval = false;
  int status = LWFB_ST_OK;
  double tstop = t0;
  bool at_boundary = true;
  int rec = 1;

  if (d.rec) write_record(d.rec);     // record 0 = raw u0

  for (;;) {
    const bool active = (status == LWFB_ST_OK) && (t < tend);
    if (!__any_sync(LFULL, active && live)) break;
    const bool can_step = active && !at_boundary;
    const bool warp_steps = __any_sync(LFULL, can_step && live);

    if (!warp_steps) {
      // ============================== warp-synchronous daily callback at t = tstop
      // soil state of the own layer (SWCHEK clamps the state itself), then all layers gathered
      double sw = y[0];
      if (sw <= 0.0) sw = 0.0; else if (sw > SWATMAX) sw = SWATMAX;
      const double w = jmin(1.0, (THSAT * sw / SWATMAX - THR) / (THSAT - THR));
      const double a = jmax(w, 1.e-6);
      const double la = lwfb_log(a);
      const double p1 = lwfb_pow_from_log(a, la, E1);
      const double psim = CPSI * lwfb_pow(p1 - 1.0, E2);
      const double p3 = lwfb_pow_from_log(a, la, E3);
      const double inner = 1.0 - lwfb_pow(1.0 - p3, 1.0 - E2);
      const double kkv = KSAT * lwfb_pow_from_log(a, la, TORT) * (inner * inner);
      const double psit = psim + PSIG;
      double psiti_a[NMAX], kk_a[NMAX], TRANI_a[NMAX], TRANIst_a[NMAX];
      for (int j = 0; j < N; j++) { psiti_a[j] = bcast(psit, j); kk_a[j] = bcast(kkv, j); }
      const double psim0 = bcast(psim, 0);
      const int day = (int)floor(t + 0.5);
      double nRNET, nSMLT, nSLVP, nTA, nRWU;
      int cbst = LWFB_ST_OK;
      DayState nds = ds;
      cb_core<NMAX>(pv, forc + (int64_t)day * LWFB_F, t, psim0, psiti_a, kk_a, nds, TRANI_a, nRNET, nSMLT, nSLVP, nTA,
                    TRANIst_a, nRWU, cbst);
      if (active) {
        y[0] = sw;
        ds = nds;
        TRANI = TRANI_a[i]; TRANIst = TRANIst_a[i]; RWUst = nRWU;
        RNET = nRNET; SMLT = nSMLT; SLVP = nSLVP; TA = nTA;
        if (cbst != LWFB_ST_OK) status = cbst;
#pragma unroll
        for (int c = 1; c < NC; c++) if (cvalid[c] && cext[c] >= 1) y[c] = 0.0;
        if (iso) {
          double dp[2];
          prec_delta(t, dp);
#pragma unroll
          for (int c = 0; c < 3; c++) { dISN[c][0] = dp[0]; dISN[c][1] = dp[1]; }
        }
        reeval = true;
        at_boundary = false;
        tstop = fmin(tstop + 1.0, tend);
      }
      continue;
    }

    // ============================== one attempted step (all lanes execute, only `can_step` lanes commit)
    // loopheader!
    double dt_s = dt;
    bool first_or_accept = (iter == 0) || accept;
    if (iter > 0 && !accept) dt_s = dt / jmin(1 / qmin, q11 / gamma);
    if (iter > 0 && accept) dt_s = dtpropose;
    double yp[NC], kfirst[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
      yp[c] = first_or_accept ? y[c] : yprev[c];
      kfirst[c] = (iter > 0 && accept && !reeval) ? kk_[c] : k1[c];
    }
    const long long iter_s = iter + 1;
    int st_s = LWFB_ST_OK;
    if (iter_s > d.maxiters) st_s = LWFB_ST_MAXITERS;
    {
      // unstable_check: NaN in any SWATI of the member
      const unsigned nanmask = __ballot_sync(LFULL, y[0] != y[0]);
      const unsigned grp = (N == 32) ? LFULL : (((1u << N) - 1u) << base);
      if (st_s == LWFB_ST_OK && (nanmask & grp)) st_s = LWFB_ST_UNSTABLE;
    }
    dt_s = dt_s < d.dtmax ? dt_s : d.dtmax;
    { double dtmin_t = fmax(eps_of(t), dtmin_opt); if (dt_s < dtmin_t) dt_s = dtmin_t; }
    if (dt_s > tstop - t) dt_s = tstop - t;
    if (st_s == LWFB_ST_OK && !(fabs(dt_s) > dtmin_opt)) st_s = LWFB_ST_DTMIN;
    const bool do_step = can_step && (st_s == LWFB_ST_OK);   // member-level: identical in all lanes of the group

    // perform_step!: rolled stage loop (s = 0 is the FSAL re-evaluation after a callback)
    double xin[NC], kcur[NC], q_[NC][5], aE[NC];
    double VV = 0, INFLI = 0, DSFLI = 0, nt = ntfli;
    long long nf_s = nf;
#pragma unroll
    for (int c = 0; c < NC; c++) { kcur[c] = kfirst[c]; aE[c] = 0; xin[c] = yp[c];
#pragma unroll
      for (int p = 0; p < 5; p++) q_[c][p] = 0; }
    const int s_first = __any_sync(LFULL, reeval && can_step) ? 0 : 1;
#pragma unroll 1
    for (int s = s_first; s <= 6; s++) {
      // ---- stage input
      if (s == 1) {
        const double a = dt_s * 0.161;
#pragma unroll
        for (int c = 0; c < NC; c++) {
          xin[c] = yp[c] + a * kfirst[c];
          aE[c] = lBT[1] * kfirst[c];
#pragma unroll
          for (int p = 0; p < 5; p++) q_[c][p] = lCQ[1][p] * kfirst[c];
        }
      } else if (s >= 2) {
#pragma unroll
        for (int c = 0; c < NC; c++) {
          xin[c] = yp[c] + dt_s * q_[c][0];
#pragma unroll
          for (int p = 0; p < 4; p++) q_[c][p] = q_[c][p + 1];
          q_[c][4] = 0;
        }
      } else {
#pragma unroll
        for (int c = 0; c < NC; c++) xin[c] = yp[c];
      }
      const double ts = (s == 0) ? t : ((s < 5) ? t + lC[s] * dt_s : t + dt_s);
      const bool skip0 = (s == 0) && !reeval;     // groups that do not need the re-evaluation keep their k1

      // ---- RHS f(x, ts)
      nf_s += skip0 ? 0 : 1;
      double sw = xin[0];
      if (sw <= 0.0) sw = 0.0; else if (sw > SWATMAX) sw = SWATMAX;
      const double gw = ext_bcast(xin, 0);
      const double w = jmin(1.0, (THSAT * sw / SWATMAX - THR) / (THSAT - THR));
      const double a = jmax(w, 1.e-6);
      const double la = lwfb_log(a);
      const double p1 = lwfb_pow_from_log(a, la, E1);
      const double psim = CPSI * lwfb_pow(p1 - 1.0, E2);
      const double theta = w * (THSAT - THR) + THR;
      const double p3 = lwfb_pow_from_log(a, la, E3);
      const double inner = 1.0 - lwfb_pow(1.0 - p3, 1.0 - E2);
      const double kkv = KSAT * lwfb_pow_from_log(a, la, TORT) * (inner * inner);
      const double lk = lwfb_log(kkv);
      const double psit = psim + PSIG;
      double BYFRAC = 0;
      if (BYPAR) {
        const double QFFC = pv.S(LWFB_S_QFFC), QFPAR = pv.S(LWFB_S_QFPAR);
        BYFRAC = jmin(1.0, lwfb_pow(QFFC, 1 - (1 / QFPAR) * (w - WETF) / (1 - WETF)));
        if (w > 0.99) BYFRAC = 1;
      }
      double SAFRAC = 0.;
      if (__any_sync(LFULL, QLAYER > 0)) {
        double sq = 0;
        for (int j = 0; j < N; j++) { const double v = bcast(sw, j); if (j < QLAYER) sq += v; }
        if (QLAYER > 0) {
          const double QFFC = pv.S(LWFB_S_QFFC), QFPAR = pv.S(LWFB_S_QFPAR);
          SAFRAC = jmin(1.0, lwfb_pow(QFFC, 1 - (1 / QFPAR) * (sq - pv.S(LWFB_S_SWATQF)) / (pv.S(LWFB_S_SWATQX) - pv.S(LWFB_S_SWATQF))));
        }
      }
      const double SRFL = jmin(1., (IMPERV + SAFRAC)) * (RNET + SMLT);
      const double SLFL = RNET + SMLT - SRFL;
      DSFLI = 0;
      if (SLOPE) {
        const double DSLOPE = pv.S(LWFB_S_DSLOPE), LEN = pv.S(LWFB_S_LENGTH_SLOPE);
        const double cs = lwfb_cos(DSLOPE), sn = lwfb_sin(DSLOPE);
        const double GRAD = cs * (psim * 2 / (1000 * LEN)) + sn * p_RHOWG;
        const double ARATIO = THICK * (1 - STONEF) * cs / (1000 * LEN);
        const double dd = kkv * ARATIO * GRAD / p_RHOWG;
        DSFLI = dd < 0 ? 0 : dd;
      }
      // first approximation of the flux to the layer below (bottom layer: drainage)
      const double psit_dn = __shfl_down_sync(LFULL, psit, 1), lk_dn = __shfl_down_sync(LFULL, lk, 1);
      double V0;
      if (bot) V0 = (DRAIN > 0.00001) ? DRAIN * kkv * (1 - STONEF) : 0.0;
      else if (fabs(psit - psit_dn) < DPSIMAX) V0 = 0;
      else {
        double kkmean = lwfb_exp((THICK_dn * lk + THICK * lk_dn) / (THICK + THICK_dn));
        kkmean = jmin(jmin(kkmean, KSAT), KSAT_dn);
        const double grad = (psit - psit_dn) / ((THICK + THICK_dn) / 2.0);
        V0 = (grad * kkmean / p_RHOWG) * (1.0 - (STONEF + STONEF_dn) / 2.0);
      }
      const double V0_up = __shfl_up_sync(LFULL, V0, 1);        // prior flux from the layer above (unused at the top)
      // ITER: min-chain over the layers in index order (the clamp to DTIMIN makes the order matter)
      const double DTRI = 1.0 - fmod(ts, 1.0);
      double DTI = jmin(DTRI, DTIMAX);
      {
        const double c1 = CITER / jmax(0.000001, fabs(nt));
        const double c2 = ((THR - theta) * THICK) / nt / 1.30;
        const unsigned negm = __ballot_sync(LFULL, nt < 0);
        double DTINEW = DTI;
        for (int j = 0; j < N; j++) {
          DTINEW = jmin(DTINEW, bcast(c1, j));
          const double c2j = bcast(c2, j);
          if ((negm >> (base + j)) & 1u) {
            DTINEW = jmin(DTINEW, c2j);
            if (DTINEW < p_DTIMIN) DTINEW = p_DTIMIN;
          }
        }
        if (DTINEW < DTI) DTI = DTINEW;
      }
      // INFLOW
      double BYFLI = SLFL * INFRAC * BYFRAC;
      INFLI = SLFL * INFRAC * (1.0 - BYFRAC);
      VV = V0;
      {
        const double room = (SWATMAX - sw) / DTI;
        // no layer limits (the common case) <=> every layer passes its test with the unmodified fluxes
        double MAXIN = room + V0 + DSFLI + TRANI;
        if (top) MAXIN = MAXIN + SLVP;
        const bool limited = top ? (INFLI > MAXIN) : (V0_up + INFLI > MAXIN);
        if (__any_sync(LFULL, limited)) {
          // bottom-up chain: layer j fixes the posterior flux from layer j-1
          for (int j = N - 1; j >= 0; j--) {
            double MX = room + VV + DSFLI + TRANI;
            double VVup = V0_up;
            if (j == 0) {
              MX = MX + SLVP;
              if (i == 0 && INFLI > MX) {
                const double too = INFLI - MX;
                INFLI = INFLI - too;
                BYFLI = BYFLI + too;
              }
            } else {
              if (i == j && V0_up + INFLI > MX) {
                if (BYFRAC > 0) {
                  if (V0_up < MX) {
                    const double too = INFLI + V0_up - MX;
                    INFLI = INFLI - too;
                    BYFLI = BYFLI + too;
                  } else {
                    VVup = MX;
                    BYFLI = BYFLI + INFLI;
                    INFLI = 0;
                  }
                } else {
                  VVup = MX - INFLI;
                }
              }
            }
            const double recv = __shfl_down_sync(LFULL, VVup, 1);   // lane j-1 receives its posterior outflow
            if (j >= 1 && i == j - 1) VV = recv;
          }
        }
      }
      const double VV_up = __shfl_up_sync(LFULL, VV, 1);
      double du0;
      if (top) du0 = INFLI - VV - DSFLI - TRANI - SLVP;
      else du0 = VV_up + INFLI - VV - DSFLI - TRANI;
      const double VN = bcast(VV, N - 1);
      double GWFL, SEEP;
      if (GSC < 1.0e-8) { SEEP = GSP * VN; GWFL = VN - SEEP; }
      else {
        SEEP = gw * GSC * GSP;
        GWFL = gw * GSC * (1. - GSP);
        if (gw - (GWFL + SEEP) * 1 < 0) { SEEP = gw / 1 * GSP; GWFL = gw / 1 * (1. - GSP); }
      }
      double sby = 0, sds = 0;
      for (int j = 0; j < N; j++) sby += bcast(BYFLI, j);
      if (__any_sync(LFULL, SLOPE)) for (int j = 0; j < N; j++) sds += bcast(DSFLI, j);
      double kn[NC];
      kn[0] = du0;
#pragma unroll
      for (int c = 1; c < NC; c++) {
        const int e = cext[c];
        double v = VN - GWFL - SEEP;                       // e == 0: GWAT
        v = (e == 1) ? SRFL + sby + sds + GWFL : v;
        v = (e == 2) ? SEEP : v;
        v = (e == 3) ? SRFL : v;
        v = (e == 4) ? SLFL : v;
        v = (e == 5) ? sby : v;
        v = (e == 6) ? sds : v;
        v = (e == 7) ? GWFL : v;
        v = (e == 8) ? VN : v;
        kn[c] = cvalid[c] ? v : 0.0;
      }
      if (!skip0) nt = du0;
      // ---- end of RHS

      if (s == 0) {
#pragma unroll
        for (int c = 0; c < NC; c++) { kfirst[c] = skip0 ? kfirst[c] : kn[c]; }
      } else {
        const int ks = s + 1;
#pragma unroll
        for (int c = 0; c < NC; c++) {
          aE[c] = aE[c] + lBT[ks] * kn[c];
#pragma unroll
          for (int p = 0; p < 5; p++) if (ks + 1 + p <= 7) q_[c][p] = q_[c][p] + lCQ[ks][p] * kn[c];
          kcur[c] = kn[c];
        }
        if (s == 6) xin[0] = sw;      // the new state u: SWCHEK clamped it in place
      }
    }
    // residuals + norm_to_use: sequential chain SWATI[0..N-1], GWAT, 8 accumulators (oracle order)
    double r2[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
      const double ut = dt_s * aE[c];
      const double r = ut / (d.abstol + jmax(fabs(yp[c]), fabs(xin[c])) * d.reltol);
      r2[c] = r * r;
    }
    double ss = 0;
    for (int j = 0; j < N; j++) ss += bcast(r2[0], j);
#pragma unroll 1
    for (int e = 0; e < 9; e++) ss += ext_bcast(r2, e);
    const double EEst = sqrt(ss / (double)(3 * N + 39));
    // loopfooter!: PI controller
    const double ttmp = t + dt_s;
    double q, q11n = q11;
    if (EEst == 0.0) q = 1 / qmax;
    else {
      q11n = fastpow(EEst, beta1);
      q = q11n / fastpow(qold, beta2);
      q = jmax(1 / qmax, jmin(1 / qmin, q / gamma));
    }
    const bool acc_s = (EEst <= 1.0);

    // ---- commit
    if (can_step) {
      // what loopheader / the checks did
      iter = iter_s;
      if (st_s != LWFB_ST_OK) status = st_s;
    }
    if (do_step) {
#pragma unroll
      for (int c = 0; c < NC; c++) { yprev[c] = yp[c]; k1[c] = kfirst[c]; kk_[c] = kcur[c]; y[c] = xin[c]; }
      dt = dt_s; nf = nf_s; ntfli = nt; q11 = q11n;
      reeval = false;
      accept = acc_s;
      if (acc_s) {
        naccept++;
        qold = jmax(EEst, qoldinit);
        const double dtnew = dt_s / q;
        tprev = t;
        t = (fabs(ttmp - tstop) < 100 * eps_of(fmax(t, tstop))) ? tstop : ttmp;
        dtpropose = d.dtmax < dtnew ? d.dtmax : dtnew;
        { double dtmin_t = fmax(eps_of(t), dtmin_opt); if (dtpropose < dtmin_t) dtpropose = dtmin_t; }
      } else {
        nreject++;
      }
    }
    const bool accepted = do_step && acc_s;
    const bool at_stop = accepted && (t == tstop);
    // ---- record + balance at the day boundary (saved before the callbacks and before this step's isotope update)
    if (__any_sync(LFULL, at_stop)) {
      double Ssum = 0;
      for (int j = 0; j < N; j++) Ssum += bcast(y[0], j);
      double ev[9];
#pragma unroll 1
      for (int e = 0; e < 9; e++) ev[e] = ext_bcast(y, e);
      if (at_stop) {
        if (d.rec) write_record(d.rec + (int64_t)rec * L * M);
        rec++;
        if (t < tend) {
          double* a = ds.acc;   // LWFBrook90R_check_balance_errors!
          const double Wt = ds.INTR + ds.INTS + ds.SNOW + Ssum + ev[0];
          const double qin = 0 + (ev[4] - ev[5]);
          const double qout = ev[8] + ev[6] + a[LWFB_A_CUM_D_TRAN] + a[LWFB_A_CUM_D_SLVP];
          const double qin_t = a[LWFB_A_CUM_D_PREC];
          const double qout_t = a[LWFB_A_CUM_D_EVAP] + ev[1] + ev[2];
          a[LWFB_A_BALERD_SWAT] = (qin - qout) - (Ssum - a[LWFB_A_STORAGESWAT]);
          a[LWFB_A_BALERD_TOTAL] = (qin_t - qout_t) - (Wt - a[LWFB_A_STORAGEWATER]);
          a[LWFB_A_STORAGESWAT] = Ssum;
          a[LWFB_A_STORAGEWATER] = Wt;
          at_boundary = true;
        }
      }
    }
    // ---- isotope advection-dispersion step with the flux caches of stage 7
    const bool do_ade = accepted && iso && (t != tstop || t < tend);
    if (__any_sync(LFULL, do_ade)) {
      double dp[2];
      prec_delta(t, dp);
      const double dtt = t - tprev;
      const double S1 = y[0];
      const double w1 = jmin(1.0, (THSAT * S1 / SWATMAX - THR) / (THSAT - THR));
      const double th1 = w1 * (THSAT - THR) + THR;
      const double Tk = TA + 273.15;
      const double Lam = pv.S(LWFB_S_DISPERSIVITY), VXYLEM = pv.S(LWFB_S_VXYLEM);
      const double ee = lwfb_exp(-535400 / (Tk * Tk) + 1393.3 / Tk + 2.1876);
      double sumT = 0;
      for (int j = 0; j < N; j++) sumT += bcast(TRANI, j);
      const double tauth = lwfb_pow(th1, 7.0 / 3) / (THSAT * THSAT);
      const double aq = Lam * fabs(VV / 1000);
      const double G1 = ext_bcast(y, 0);
      const double VN = bcast(VV, N - 1);
      const double VV_up = __shfl_up_sync(LFULL, VV, 1);
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const double R = s == 0 ? R18 : R2H;
        const double D0 = (s == 0 ? 0.96691 : 0.98331) * 1e-9 * ee * 3600 * 24;
        const double Cxy = d_to_x(dXY[s], R), Cinf = d_to_x(dp[s], R);
        const double Dd = D0 * tauth * th1 + aq;
        const double C = d_to_x(dS[s], R);
        const double Ctr = (TRANI < 0) ? Cxy : C;
        const double C_up = __shfl_up_sync(LFULL, C, 1), C_dn = __shfl_down_sync(LFULL, C, 1);
        const double Dd_up = __shfl_up_sync(LFULL, Dd, 1), Dd_dn = __shfl_down_sync(LFULL, Dd, 1);
        const double diff_upp = top ? 0.0 : ((Dd_up + Dd) / 2) * ((C - C_up) / DZ_up);
        const double diff_low = bot ? 0.0 : ((Dd + Dd_dn) / 2) * ((C_dn - C) / DZ_dn);
        const double qC_upp = top ? 0.0 : VV_up * C_up;
        const double qC_low = VV * C;
        const double slv = top ? C : 0.0;
        double innr = -diff_upp * 1000 + diff_low * 1000 + qC_upp;
        innr = innr - qC_low;
        innr = innr + INFLI * Cinf;
        innr = innr - DSFLI * C;
        innr = innr - SLVP * slv;
        innr = innr - TRANI * Ctr;
        const double dC = -C / S1 * ntfli + 1 / S1 * innr;
        const double dd = dxdt_to_dddt(dC, C, R);
        const double CN = bcast(C, N - 1);
        double ddG;
        if ((G1 == 0) && (VN == 0)) ddG = NAN;
        else if ((G1 == 0) && (VN > 0)) ddG = 0;
        else {
          const double CG = d_to_x(dG[s], R);
          const double dCG = VN / G1 * (CN - CG);
          ddG = dxdt_to_dddt(dCG, CG, R);
        }
        double swt = 0;
        const double ctt = Ctr * TRANI;
        for (int j = 0; j < N; j++) swt += bcast(ctt, j);
        const double dRW = x_to_d(swt / sumT, R);
        const double ddxy = (dRW != dRW) ? 0.0 : sumT / VXYLEM * (dRW - dXY[s]);
        if (do_ade) {
          dG[s] += dtt * ddG;
          dS[s] += dtt * dd;
          dRWU[s] = dRW;
          dXY[s] += dtt * ddxy;
        }
      }
    }
  }
  if (live) {
    if (d.rec) {
      for (; rec <= d.n_days; rec++) {
        double* dst = d.rec + (int64_t)rec * L * M;
        for (int j = i; j < L; j += N) dst[(int64_t)j * M + m] = NAN;
      }
    }
    write_record(d.fin);
    if (top) {
      d.status[m] = status;
      d.counters[(int64_t)LWFB_C_NACCEPT * M + m] = naccept;
      d.counters[(int64_t)LWFB_C_NREJECT * M + m] = nreject;
      d.counters[(int64_t)LWFB_C_NF * M + m] = nf;
    }
  }
}

#ifdef __CUDACC__
template <int NMAX, int EMAX>
__global__ void __launch_bounds__(128) integrate_lanes(DevData d) {
  lanes_body<NMAX, EMAX>(d, threadIdx.x & 31, ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
}
#endif

}  // namespace lwfb
