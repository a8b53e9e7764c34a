// lwfb90_member.cuh — thread-per-member integrator (runtime layer count).
//
// The fallback for layer counts the lane-per-layer kernel (lwfb90_lanes.cuh) does not cover
// (N < 3 or N > 32), and the second, independently written CUDA implementation used to cross-check it.
// One member per thread; state in thread-local storage.
#pragma once
#include "lwfb90_device.cuh"
#include "lwfb90_tsit5.cuh"

namespace lwfb {

template <int NT>
__device__ void integrate_member(const DevData& d, const int64_t m) {
  constexpr int NMAX = Member<NT>::NMAX;
  constexpr int NEMAX = NMAX + 9;
  Member<NT> mb;
  mb.load(d, m);
  const int N = mb.N, NE = N + 9, L = d.L;
  if (d.ntfli0) for (int i = 0; i < N; i++) mb.NTFLI[i] = __ldg(d.ntfli0 + m * N + i);   // warm du_NTFLI cache
  const bool iso = d.iso != 0;
  const SIdx ix(N, d.iso);
  const double* __restrict__ forc = d.forc + (int64_t)__ldg(d.site + m) * d.n_rows * LWFB_F;
  auto U0 = [&](int j) { return __ldg(d.u0 + m * L + j); };

  // ---- evolving state y = (SWATI[N], GWAT, 8 integrated accumulators) and the rest of u
  double y[NEMAX], yprev[NEMAX], k[7][NEMAX];
  double TRANIst[NMAX], RWUst, XYmm;
  DayState ds;
  IsoState<NT> is;
  double dISN[3][2];  // delta of INTS, INTR, SNOW
  for (int i = 0; i < N; i++) { y[i] = U0(ix.SWATI + i); TRANIst[i] = U0(ix.TRANI + i); }
  y[N] = U0(ix.GWAT);
  for (int a = 0; a < LWFB_NACCUM; a++) ds.acc[a] = U0(ix.ACC + a);
  for (int q = 0; q < 8; q++) y[N + 1 + q] = ds.acc[kAccPos[q]];
  ds.INTS = U0(ix.INTS); ds.INTR = U0(ix.INTR); ds.SNOW = U0(ix.SNOW); ds.CC = U0(ix.CC); ds.SNOWLQ = U0(ix.SNOWLQ);
  RWUst = U0(ix.RWU); XYmm = U0(ix.XYLEM);
  if (iso) {
    for (int s = 0; s < 2; s++) {
      for (int i = 0; i < N; i++) is.dS[s][i] = U0(ix.SWATI + (s + 1) * N + i);
      is.dG[s] = U0(ix.GWAT + 1 + s); is.dRWU[s] = U0(ix.RWU + 1 + s); is.dXY[s] = U0(ix.XYLEM + 1 + s);
      dISN[0][s] = U0(ix.INTS + 1 + s); dISN[1][s] = U0(ix.INTR + 1 + s); dISN[2][s] = U0(ix.SNOW + 1 + s);
    }
  }

  // writes the full reference-order state record (src/func_DiffEq_definition_u0.jl:42-73)
  auto write_record = [&](double* __restrict__ dst) {
    auto W = [&](int j, double v) { dst[j] = v; };
    for (int i = 0; i < N; i++) W(ix.SWATI + i, y[i]);
    W(ix.GWAT, y[N]); W(ix.INTS, ds.INTS); W(ix.INTR, ds.INTR); W(ix.SNOW, ds.SNOW);
    W(ix.RWU, RWUst); W(ix.XYLEM, XYmm); W(ix.CC, ds.CC); W(ix.SNOWLQ, ds.SNOWLQ);
    for (int i = 0; i < N; i++) W(ix.TRANI + i, TRANIst[i]);
    if (iso) {
      for (int s = 0; s < 2; s++) {
        for (int i = 0; i < N; i++) {
          W(ix.SWATI + (s + 1) * N + i, is.dS[s][i]);
          W(ix.TRANI + (s + 1) * N + i, U0(ix.TRANI + (s + 1) * N + i));  // never updated by the reference
        }
        W(ix.GWAT + 1 + s, is.dG[s]); W(ix.RWU + 1 + s, is.dRWU[s]); W(ix.XYLEM + 1 + s, is.dXY[s]);
        W(ix.INTS + 1 + s, dISN[0][s]); W(ix.INTR + 1 + s, dISN[1][s]); W(ix.SNOW + 1 + s, dISN[2][s]);
      }
    }
    for (int j = 0; j < 3 * N; j++) W(ix.AUX + j, U0(ix.AUX + j));  // aux is never written after init
    for (int a = 0; a < LWFB_NACCUM; a++) W(ix.ACC + a, ds.acc[a]);
    for (int q = 0; q < 8; q++) W(ix.ACC + kAccPos[q], y[N + 1 + q]);
  };
  // goodness-of-fit sums at the observation days (lwfb_set_observations), from what daily record r holds
  auto score_rec = [&](int r, bool reached) {
    if (d.n_obs_layers == 0) return;
    const int ko = __ldg(d.obs_of_rec + r);
    if (ko < 0) return;
    double* a = d.sacc + m * (2 * LWFB_SA);
    for (int l = 0; l < d.n_obs_layers; l++) {
      const int il = __ldg(d.obs_layer + l);
      const double sw = reached ? (r == 0 ? U0(ix.SWATI + il) : y[il]) : NAN;
      if (d.obs_theta)
        score_add(a, reached ? score_theta(mb.THSAT[il], mb.THR[il], mb.SWATMAX[il], sw) : NAN,
                  __ldg(d.obs_theta + (int64_t)ko * d.n_obs_layers + l));
      if (d.obs_d18o && iso)
        score_add(a + LWFB_SA, reached ? (r == 0 ? U0(ix.SWATI + N + il) : is.dS[0][il]) : NAN,
                  __ldg(d.obs_d18o + (int64_t)ko * d.n_obs_layers + l));
    }
  };
  // delta of precipitation at time t: the value of forcing row r holds for t in (r-0.5, r+0.5]
  auto prec_delta = [&](double t, double* dp) {
    int r = (int)ceil(t - 0.5);
    r = r < 0 ? 0 : (r > d.n_rows - 1 ? d.n_rows - 1 : r);
    dp[0] = __ldg(forc + (int64_t)r * LWFB_F + LWFB_F_D18O);
    dp[1] = __ldg(forc + (int64_t)r * LWFB_F + LWFB_F_D2H);
  };
  auto cb_deltas = [&](double t) {  // src/func_DiffEq_definition_cb.jl:304-530 (live :380-387, :520-525)
    if (!iso) return;
    double dp[2];
    prec_delta(t, dp);
    for (int c = 0; c < 3; c++) { dISN[c][0] = dp[0]; dISN[c][1] = dp[1]; }
  };

  const double qmin = 0.2, qmax = 10.0, gamma = 0.9, beta1 = 0.14, beta2 = 0.08, qoldinit = 1e-4;
  const double t0 = (double)d.t0_day, tend = (double)(d.t0_day + d.n_days);
  double t = t0, tprev = t0, dt = d.dt_init, dtpropose = d.dt_init, qold = qoldinit, q11 = 1.0;
  const double dtmin_opt = fmax(eps_of(tend), eps_of(t0));
  long long naccept = 0, nreject = 0, iter = 0;
  bool accept = false, reeval = false;
  int status = LWFB_ST_OK;

  // ---- init: save u0, initial_affect of the daily callbacks, FSAL
  double* const rec_m = d.rec ? d.rec + m * (int64_t)(d.n_days + 1) * L : nullptr;   // [n_days+1][L] of this member
  if (d.rec) write_record(rec_m);
  score_rec(0, true);
  cb_amounts<NT>(mb, forc + (int64_t)d.t0_day * LWFB_F, t0, y, ds, TRANIst, RWUst);
  for (int q = 0; q < 8; q++) y[N + 1 + q] = 0.0;
  cb_deltas(t0);
  for (int j = 0; j < NE; j++) yprev[j] = y[j];
  mb.rhs(yprev, yprev[N], t, k[0]);
  double tstop = fmin(t0 + 1.0, tend);
  int rec = 1;
  status = mb.status;   // a failing initial callback stops the member before its first step

  while (status == LWFB_ST_OK && t < tend) {
    // loopheader!
    if (iter > 0) {
      if (accept) {
        for (int j = 0; j < NE; j++) yprev[j] = y[j];
        dt = dtpropose;
        if (reeval) mb.rhs(y, y[N], t, k[0]);
        else for (int j = 0; j < NE; j++) k[0][j] = k[6][j];
      } else {
        dt = dt / jmin(1 / qmin, q11 / gamma);
      }
    }
    iter++;
    if (iter > d.maxiters) { status = LWFB_ST_MAXITERS; break; }
    {
      bool bad = false;
      for (int i = 0; i < N; i++) bad |= (y[i] != y[i]);
      if (bad) { status = LWFB_ST_UNSTABLE; break; }
    }
    dt = dt < d.dtmax ? dt : d.dtmax;
    { double dtmin_t = fmax(eps_of(t), dtmin_opt); if (dt < dtmin_t) dt = dtmin_t; }
    if (dt > tstop - t) dt = tstop - t;
    if (!(fabs(dt) > dtmin_opt)) { status = LWFB_ST_DTMIN; break; }
    // perform_step! Tsit5; only SWATI and GWAT feed the RHS, the accumulators ride along.
    // The stage loop is kept rolled so that the RHS is instantiated once.
    double tmp[NEMAX];
#pragma unroll 1
    for (int s = 1; s <= 5; s++) {
      if (s == 1) {
        const double a = dt * a21;
        for (int j = 0; j <= N; j++) tmp[j] = yprev[j] + a * k[0][j];
      } else {
        for (int j = 0; j <= N; j++) {
          double acc = kA[s][0] * k[0][j];
          for (int l = 1; l < s; l++) acc = acc + kA[s][l] * k[l][j];
          tmp[j] = yprev[j] + dt * acc;
        }
      }
      mb.rhs(tmp, tmp[N], (s < 5) ? t + kC[s] * dt : t + dt, k[s]);
    }
    for (int j = 0; j < NE; j++)
      y[j] = yprev[j] + dt * (a71 * k[0][j] + a72 * k[1][j] + a73 * k[2][j] + a74 * k[3][j] + a75 * k[4][j] + a76 * k[5][j]);
    mb.rhs(y, y[N], t + dt, k[6]);  // clamps the new SWATI in place
    // residuals + norm_to_use (src/func_DiffEq_definition_ode.jl:20-37): RMS over 3N+39 entries
    double ss = 0;
    for (int j = 0; j < NE; j++) {
      double ut = dt * (bt1 * k[0][j] + bt2 * k[1][j] + bt3 * k[2][j] + bt4 * k[3][j] + bt5 * k[4][j] + bt6 * k[5][j] + bt7 * k[6][j]);
      double r = ut / (d.abstol + jmax(fabs(yprev[j]), fabs(y[j])) * d.reltol);
      ss += r * r;
    }
    double EEst = sqrt(ss / (double)(3 * N + 39));
    // loopfooter!: PI controller
    reeval = false;
    double ttmp = t + dt;
    double q;
    if (EEst == 0.0) q = 1 / qmax;
    else {
      q11 = fastpow(EEst, beta1);
      q = q11 / fastpow(qold, beta2);
      q = jmax(1 / qmax, jmin(1 / qmin, q / gamma));
    }
    accept = (EEst <= 1.0);
    if (accept) {
      naccept++;
      qold = jmax(EEst, qoldinit);
      double dtnew = dt / q;
      tprev = t;
      t = (fabs(ttmp - tstop) < 100 * eps_of(fmax(t, tstop))) ? tstop : ttmp;
      dtpropose = d.dtmax < dtnew ? d.dtmax : dtnew;
      { double dtmin_t = fmax(eps_of(t), dtmin_opt); if (dtpropose < dtmin_t) dtpropose = dtmin_t; }
      double dp[2] = {0, 0};
      if (iso) prec_delta(t, dp);
      if (t == tstop) {
        // the record is saved before the callbacks of this instant (and before this step's ADE update)
        if (d.rec) write_record(rec_m + (int64_t)rec * L);
        score_rec(rec, true);
        rec++;
        if (t < tend) {
          // LWFBrook90R_check_balance_errors! src/func_DiffEq_definition_cb.jl:1255-1314
          {
            double* a = ds.acc;
            double Ssum = 0;
            for (int i = 0; i < N; i++) Ssum += y[i];
            double Wt = ds.INTR + ds.INTS + ds.SNOW + Ssum + y[N];
            double qin = 0 + (y[N + 4] - y[N + 5]);
            double qout = y[N + 8] + y[N + 6] + a[LWFB_A_CUM_D_TRAN] + a[LWFB_A_CUM_D_SLVP];
            double qin_t = a[LWFB_A_CUM_D_PREC];
            double qout_t = a[LWFB_A_CUM_D_EVAP] + y[N + 1] + y[N + 2];
            a[LWFB_A_BALERD_SWAT] = (qin - qout) - (Ssum - a[LWFB_A_STORAGESWAT]);
            a[LWFB_A_BALERD_TOTAL] = (qin_t - qout_t) - (Wt - a[LWFB_A_STORAGEWATER]);
            a[LWFB_A_STORAGESWAT] = Ssum;
            a[LWFB_A_STORAGEWATER] = Wt;
          }
          if (iso) ade_step<NT>(mb, is, y, y[N], t - tprev, dp);
          int day = (int)floor(t + 0.5);
          cb_amounts<NT>(mb, forc + (int64_t)day * LWFB_F, t, y, ds, TRANIst, RWUst);
          for (int q8 = 0; q8 < 8; q8++) y[N + 1 + q8] = 0.0;
          cb_deltas(t);
          reeval = true;
          tstop = fmin(tstop + 1.0, tend);
        }
      } else {
        if (iso) ade_step<NT>(mb, is, y, y[N], t - tprev, dp);
      }
    } else {
      nreject++;
    }
    if (mb.status != LWFB_ST_OK) { status = mb.status; break; }
  }
  for (int r = rec; r <= d.n_days; r++) score_rec(r, false);
  if (d.rec) {
    for (; rec <= d.n_days; rec++)
      for (int j = 0; j < L; j++) rec_m[(int64_t)rec * L + j] = NAN;
  }
  write_record(d.fin + m * L);
  if (d.ntfli_out) for (int i = 0; i < N; i++) d.ntfli_out[m * N + i] = mb.NTFLI[i];
  d.status[m] = status;
  d.counters[m * LWFB_NCOUNTERS + LWFB_C_NACCEPT] = naccept;
  d.counters[m * LWFB_NCOUNTERS + LWFB_C_NREJECT] = nreject;
  d.counters[m * LWFB_NCOUNTERS + LWFB_C_NF] = mb.nf;
}

}  // namespace lwfb
