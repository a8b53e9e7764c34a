// lwfb90_wide.cuh — the integrator for deep profiles (32 < N <= LWFB_MAX_LAYERS): one WARP per member.
//
// The lane-per-layer kernel (lwfb90_lanes.cuh) keeps a layer in the registers of one lane, which ends at 32 layers.
// Here a member owns a warp and a block of shared memory: every per-layer quantity - run constants, the Runge-Kutta
// state with its seven stage derivatives, the flux caches the reference keeps in `p`, the isotope state, the work
// arrays of one RHS evaluation - is an array [N + 9] in shared memory, and lane l works on layers l, l + 32, l + 64, ...
// (the refined grids of the reference's integration tests: INFEXP1 N = 42, the G4 golden N = 40,
// Hammel_sand-NLayer-400, test/02-integration-tests.jl:383).  Couplings between layers are plain shared-memory reads
// after a __syncwarp:
//   * Darcy flux, advection-dispersion: neighbour values                       -> a[i + 1], a[i - 1]
//   * sums in layer order (error norm, bypass / downslope totals, storage)     -> every lane adds the N terms in index
//     order from shared memory (broadcast reads): the oracle's summation order, the same value in all lanes
//   * INFLOW's bottom-up oversaturation limiter, ITER's clamped min-chain       -> a warp vote finds out whether any layer
//     limits / any term needs the clamp (rare); only then lane 0 walks the chain, otherwise the posterior fluxes are the
//     prior ones and the chain is a warp min-reduction
// Member-level scalars (clock, step size, controller, GWAT's delta ...) are computed redundantly by all lanes from
// identical inputs.  The daily callback is a scalar state machine of ~2000 operations per member-day against
// ~150 x N x 400 of the RHS evaluations of that day: lane 0 runs it (cb_core, with its per-layer work arrays in the
// member's shared memory - LWFB_CB_SCRATCH) after all lanes have computed the soil state of their layers.
//
// Arithmetic: the formulas, operation order and elementary functions of the thread-per-member kernel
// (lwfb90_device.cuh Member::rhs, lwfb90_member.cuh), i.e. of the CPU oracle: results are bit-identical to both.
// Reference: f_LWFBrook90! src/func_DiffEq_definition_f.jl:7-243, MSBITERATE! src/func_MSB_functions.jl:565-697,
// INFLOW! src/module_WAT.jl:313-457, ITER :459-571, the isotope step src/func_DiffEq_definition_cb.jl:635-1250.
#pragma once
#ifndef LWFB_CB_SCRATCH
#error "lwfb90_wide.cuh needs the workspace form of the daily callback: compile its translation unit with -DLWFB_CB_SCRATCH"
#endif
#include "lwfb90_device.cuh"
#include "lwfb90_tsit5.cuh"

namespace lwfb {

#ifndef LFULL
#define LFULL 0xffffffffu
#endif

// rows of a member's shared-memory block: row k is a[k * NS .. k * NS + NS), NS = N + 9
enum {
  // run constants of the layers
  WA_THSAT = 0, WA_THR, WA_SWATMAX, WA_KSAT, WA_TORT, WA_PSIG, WA_THICK, WA_STONEF, WA_INFRAC, WA_CPSI,
  WA_E1, WA_E2, WA_E3, WA_E4,
  // Runge-Kutta: state at the start of the step, stage input / new state, seven stage derivatives; entries
  // [0, N) SWATI, N GWAT, N + 1 .. N + 8 the integrated accumulators
  WA_Y, WA_X, WA_K0, WA_K6 = WA_K0 + 6,
  // caches of the RHS that outlive the call (src/func_DiffEq_definition_p.jl:503-581)
  WA_TRANI, WA_NTFLI, WA_VRFLI, WA_INFLI, WA_DSFLI,
  // delta18O / delta2H of SWATI, u.TRANI.mmday
  WA_DS18, WA_DS2H, WA_TRANIST,
  // work arrays: total potential and conductivity of the layers (RHS, daily callback), then seven rows that the RHS,
  // the isotope step and the daily callback (as its 7 N workspace) use in turn
  WA_PSITI, WA_KK, WA_T0, WA_T6 = WA_T0 + 6,
  WA_COUNT
};
#define LWFB_WIDE_DOUBLES_PER_MEMBER(N) ((size_t)lwfb::WA_COUNT * ((N) + 9))

struct Wide {
  int N, NS, lane;
  double* a;
  ParamView pv;
  double DPSIMAX, DSWMAX, DTIMAX, DRAIN, GSC, GSP, IMPERV, QFFC, QFPAR, SWATQX, SWATQF, DSLOPE, LENGTH_SLOPE;
  int QLAYER, BYPAR;
  double RS, SLVP, TA;       // RNET + SMLT, SLVP, TA of the day (set by the daily callback)
  long long nf;

  __device__ __forceinline__ double* A(int k) const { return a + (size_t)k * NS; }

  __device__ void load(const DevData& d, int64_t m, int lane_, double* smem) {
    N = d.N; NS = N + 9; lane = lane_; a = smem;
    pv.N = N; pv.m = m; pv.M = d.M; pv.scal = d.scal; pv.lay = d.lay; pv.scratch = A(WA_T0);
    for (int i = lane; i < N; i += 32) {
      A(WA_THSAT)[i] = pv.Lp(LWFB_L_THSAT, i); A(WA_THR)[i] = pv.Lp(LWFB_L_THR, i); A(WA_SWATMAX)[i] = pv.Lp(LWFB_L_SWATMAX, i);
      A(WA_KSAT)[i] = pv.Lp(LWFB_L_KSAT, i); A(WA_TORT)[i] = pv.Lp(LWFB_L_TORT, i); A(WA_PSIG)[i] = pv.Lp(LWFB_L_PSIG, i);
      A(WA_THICK)[i] = pv.Lp(LWFB_L_THICK, i); A(WA_STONEF)[i] = pv.Lp(LWFB_L_STONEF, i); A(WA_INFRAC)[i] = pv.Lp(LWFB_L_INFRAC, i);
      const double n = pv.Lp(LWFB_L_NPAR, i), mp = pv.Lp(LWFB_L_MPAR, i);
      A(WA_CPSI)[i] = 9.81 * (-1.0 / pv.Lp(LWFB_L_ALPHA, i));
      A(WA_E1)[i] = -1.0 / mp; A(WA_E2)[i] = 1.0 / n; A(WA_E3)[i] = n / (n - 1.0); A(WA_E4)[i] = 1.0 - 1.0 / n;
      A(WA_TRANI)[i] = NAN; A(WA_NTFLI)[i] = NAN; A(WA_VRFLI)[i] = NAN; A(WA_INFLI)[i] = NAN; A(WA_DSFLI)[i] = NAN;
    }
    DPSIMAX = pv.S(LWFB_S_DPSIMAX); DSWMAX = pv.S(LWFB_S_DSWMAX); DTIMAX = pv.S(LWFB_S_DTIMAX); DRAIN = pv.S(LWFB_S_DRAIN);
    GSC = pv.S(LWFB_S_GSC); GSP = pv.S(LWFB_S_GSP); IMPERV = pv.S(LWFB_S_IMPERV); QFFC = pv.S(LWFB_S_QFFC);
    QFPAR = pv.S(LWFB_S_QFPAR); SWATQX = pv.S(LWFB_S_SWATQX); SWATQF = pv.S(LWFB_S_SWATQF); DSLOPE = pv.S(LWFB_S_DSLOPE);
    LENGTH_SLOPE = pv.S(LWFB_S_LENGTH_SLOPE);
    QLAYER = (int)pv.S(LWFB_S_QLAYER); BYPAR = (pv.S(LWFB_S_BYPAR) == 1.0) ? 1 : 0;
    RS = SLVP = TA = NAN;
    nf = 0;
    __syncwarp(LFULL);
  }

  // SWCHEK! + derive_auxiliary_SOILVAR! for one layer (src/module_KPT.jl:761-790, :445-464); sw is clamped in place
  __device__ __forceinline__ void soilvar1(int i, double* sw, double& W, double& psim, double& theta, double& kk) const {
    double s = sw[i];
    if (s <= 0.0) s = 0.0; else if (s > A(WA_SWATMAX)[i]) s = A(WA_SWATMAX)[i];
    sw[i] = s;
    const double THSAT = A(WA_THSAT)[i], THR = A(WA_THR)[i];
    const double w = jmin(1.0, div_c(div_c(THSAT * s, A(WA_SWATMAX)[i]) - THR, THSAT - THR));
    W = w;
    const double aw = jmax(w, 1.e-6);
    const double la = m_log(aw);
    const double p1 = m_exp_of_scaled_log(A(WA_E1)[i], la, aw);
    psim = A(WA_CPSI)[i] * m_pow(p1 - 1.0, A(WA_E2)[i]);
    theta = w * (THSAT - THR) + THR;
    const double p3 = m_exp_of_scaled_log(A(WA_E3)[i], la, aw);
    const double inner = 1.0 - m_pow(1.0 - p3, A(WA_E4)[i]);
    kk = A(WA_KSAT)[i] * m_exp_of_scaled_log(A(WA_TORT)[i], la, aw) * (inner * inner);
  }

  // sum of a row in layer order, the same value in every lane (row complete and visible: caller has synchronised)
  __device__ __forceinline__ double ordered_sum(const double* v, int n) const {
    double s = 0;
    for (int i = 0; i < n; i++) s += *(const volatile double*)(v + i);
    return s;
  }

  // INFLOW! src/module_WAT.jl:313-457.  BYFRAC, V0: rows T1, T2 of this evaluation; BYFLI: row T3.
  __device__ __forceinline__ void inflow(double DTI, double SLFL, const double* sw) {
    const double* BYFRAC = A(WA_T0 + 1);
    const double* V0 = A(WA_T0 + 2);
    double* BYFLI = A(WA_T0 + 3);
    double *INFLI = A(WA_INFLI), *VRFLI = A(WA_VRFLI);
    const double *DSFLI = A(WA_DSFLI), *TRANI = A(WA_TRANI), *SWATMAX = A(WA_SWATMAX), *INFRAC = A(WA_INFRAC);
    bool limited = false;
    for (int i = lane; i < N; i += 32) {
      BYFLI[i] = SLFL * INFRAC[i] * BYFRAC[i];
      INFLI[i] = SLFL * INFRAC[i] * (1.0 - BYFRAC[i]);
      VRFLI[i] = V0[i];
      // does the layer pass its test with the unmodified fluxes?  (if all do, the chain below changes nothing)
      double MAXIN = (SWATMAX[i] - sw[i]) / DTI + V0[i] + DSFLI[i] + TRANI[i];
      if (i == 0) { MAXIN = MAXIN + SLVP; limited = limited || (INFLI[0] > MAXIN); }
      else limited = limited || (V0[i - 1] + INFLI[i] > MAXIN);
    }
    if (!__any_sync(LFULL, limited)) { __syncwarp(LFULL); return; }
    __syncwarp(LFULL);
    if (lane == 0) {
      for (int i = N - 1; i >= 0; i--) {
        double MAXIN = (SWATMAX[i] - sw[i]) / DTI + VRFLI[i] + DSFLI[i] + TRANI[i];
        if (i == 0) {
          MAXIN = MAXIN + SLVP;
          if (INFLI[0] > MAXIN) {
            const double too = INFLI[0] - MAXIN;
            INFLI[0] = INFLI[0] - too;
            BYFLI[0] = BYFLI[0] + too;
          }
        } else {
          const double vp = V0[i - 1];
          if (vp + INFLI[i] > MAXIN) {
            if (BYFRAC[i] > 0) {
              if (vp < MAXIN) {
                VRFLI[i - 1] = vp;
                const double too = INFLI[i] + vp - MAXIN;
                INFLI[i] = INFLI[i] - too;
                BYFLI[i] = BYFLI[i] + too;
              } else {
                VRFLI[i - 1] = MAXIN;
                BYFLI[i] = BYFLI[i] + INFLI[i];
                INFLI[i] = 0;
              }
            } else {
              VRFLI[i - 1] = MAXIN - INFLI[i];
            }
          } else {
            VRFLI[i - 1] = vp;
          }
        }
      }
    }
    __syncwarp(LFULL);
  }

  // f(sw, gwat, t) -> du (rows of N + 9 entries); sw is clamped in place.  Called by the whole warp.
  __device__ __noinline__ void rhs(double* sw, double t, double* du) {
    nf++;
    double *psiti = A(WA_PSITI), *kk = A(WA_KK), *lk = A(WA_T0), *BYFRAC = A(WA_T0 + 1), *V0 = A(WA_T0 + 2),
           *BYFLI = A(WA_T0 + 3), *C1 = A(WA_T0 + 4), *C2 = A(WA_T0 + 5);
    double *NTFLI = A(WA_NTFLI), *DSFLI = A(WA_DSFLI), *VRFLI = A(WA_VRFLI), *INFLI = A(WA_INFLI);
    const double *THICK = A(WA_THICK), *STONEF = A(WA_STONEF), *KSAT = A(WA_KSAT), *TRANI = A(WA_TRANI);
    const double gwat = *(const volatile double*)(sw + N);
    const bool slope = !(LENGTH_SLOPE == 0 || DSLOPE == 0);
    const double cs = slope ? lwfb_cos(DSLOPE) : 1.0, sn = slope ? lwfb_sin(DSLOPE) : 0.0;
    bool slow = false;               // ITER: some term is NaN or below DTIMIN (the clamp makes the order matter)
    double vmin = INFINITY;
    for (int i = lane; i < N; i += 32) {
      double W, psim, theta, kki;
      soilvar1(i, sw, W, psim, theta, kki);
      kk[i] = kki;
      psiti[i] = psim + A(WA_PSIG)[i];
      lk[i] = m_log(kki);
      if (BYPAR) {
        const double wf = pv.Lp(LWFB_L_WETF, i);
        double bf = jmin(1.0, m_pow(QFFC, 1 - (1 / QFPAR) * (W - wf) / (1 - wf)));
        if (W > 0.99) bf = 1;
        BYFRAC[i] = bf;
      } else BYFRAC[i] = 0;
      if (!slope) DSFLI[i] = 0;
      else {
        const double GRAD = cs * (psim * 2 / (1000 * LENGTH_SLOPE)) + sn * p_RHOWG;
        const double ARATIO = THICK[i] * (1 - STONEF[i]) * cs / (1000 * LENGTH_SLOPE);
        const double dd = kki * ARATIO * GRAD / p_RHOWG;
        DSFLI[i] = dd < 0 ? 0 : dd;
      }
      // ITER's two bounds of the layer, from the net flow the previous call left behind
      const double nt = NTFLI[i];
      const double c1 = 0.01 * DSWMAX * THICK[i] * A(WA_THSAT)[i] * (1 - STONEF[i]) / jmax(0.000001, fabs(nt));
      const double avail = (A(WA_THR)[i] - theta) * THICK[i];
      const bool neg = nt < 0;
      const double c2 = neg ? div_c(avail / nt, 1.30) : 0.0;
      C1[i] = c1; C2[i] = c2;
      const double v = neg ? (c2 < c1 ? c2 : c1) : c1;
      slow = slow || (c1 != c1) || (neg && (c2 != c2)) || !(v >= p_DTIMIN);
      vmin = v < vmin ? v : vmin;
    }
    __syncwarp(LFULL);
    double SAFRAC = 0.;
    if (QLAYER > 0) {
      const double s = ordered_sum(sw, QLAYER);
      SAFRAC = jmin(1.0, m_pow(QFFC, 1 - (1 / QFPAR) * (s - SWATQF) / (SWATQX - SWATQF)));
    }
    const double SRFL = jmin(1., (IMPERV + SAFRAC)) * RS;
    const double SLFL = RS - SRFL;
    // first approximation of the vertical flow
    for (int i = lane; i < N; i += 32) {
      if (i == N - 1) V0[i] = (DRAIN > 0.00001) ? DRAIN * kk[i] * (1 - STONEF[i]) : 0.0;
      else if (fabs(psiti[i] - psiti[i + 1]) < DPSIMAX) V0[i] = 0;
      else {
        double kkmean = m_exp(div_c(THICK[i + 1] * lk[i] + THICK[i] * lk[i + 1], THICK[i] + THICK[i + 1]));
        kkmean = jmin(jmin(kkmean, KSAT[i]), KSAT[i + 1]);
        const double grad = div_c(psiti[i] - psiti[i + 1], (THICK[i] + THICK[i + 1]) / 2.0);
        V0[i] = div_c(grad * kkmean, p_RHOWG) * (1.0 - (STONEF[i] + STONEF[i + 1]) / 2.0);
      }
    }
    __syncwarp(LFULL);
    const double DTRI = 1.0 - fmod(t, 1.0);
    double DTI = jmin(DTRI, DTIMAX);
    inflow(DTI, SLFL, sw);
    // ITER (MvG branch)
    double DTINEW = DTI;
    if (__any_sync(LFULL, slow) || !(DTI >= p_DTIMIN)) {
      for (int i = 0; i < N; i++) {
        DTINEW = jmin(DTINEW, *(const volatile double*)(C1 + i));
        if (*(const volatile double*)(NTFLI + i) < 0) {
          DTINEW = jmin(DTINEW, *(const volatile double*)(C2 + i));
          if (DTINEW < p_DTIMIN) DTINEW = p_DTIMIN;
        }
      }
    } else {
      // no NaN, nothing below DTIMIN: the chain is a plain minimum
      for (int o = 16; o > 0; o >>= 1) { const double w = __shfl_xor_sync(LFULL, vmin, o); vmin = w < vmin ? w : vmin; }
      if (vmin < DTINEW) DTINEW = vmin;
    }
    if (DTINEW < DTI) {
      DTI = DTINEW;
      inflow(DTI, SLFL, sw);
    }
    __syncwarp(LFULL);               // every lane has read the old NTFLI
    bool anyby = false;
    for (int i = lane; i < N; i += 32) {
      double ntf;
      if (i == 0) ntf = INFLI[i] - VRFLI[i] - DSFLI[i] - TRANI[i] - SLVP;
      else ntf = VRFLI[i - 1] + INFLI[i] - VRFLI[i] - DSFLI[i] - TRANI[i];
      NTFLI[i] = ntf;
      du[i] = ntf;
      anyby = anyby || (BYFLI[i] != 0.0);
    }
    anyby = __any_sync(LFULL, anyby);
    __syncwarp(LFULL);
    double GWFL, SEEP;
    const double VN = *(const volatile double*)(VRFLI + N - 1);
    if (GSC < 1.0e-8) { SEEP = GSP * VN; GWFL = VN - SEEP; }
    else {
      SEEP = gwat * GSC * GSP;
      GWFL = gwat * GSC * (1. - GSP);
      if (gwat - (GWFL + SEEP) * 1 < 0) { SEEP = gwat / 1 * GSP; GWFL = gwat / 1 * (1. - GSP); }
    }
    const double sby = anyby ? ordered_sum(BYFLI, N) : 0.0;      // (a sum of zeros of either sign, started at +0, is +0)
    const double sds = slope ? ordered_sum(DSFLI, N) : 0.0;
    if (lane == 0) {
      du[N + 0] = VN - GWFL - SEEP;
      du[N + 1] = SRFL + sby + sds + GWFL;
      du[N + 2] = SEEP;
      du[N + 3] = SRFL;
      du[N + 4] = SLFL;
      du[N + 5] = sby;
      du[N + 6] = sds;
      du[N + 7] = GWFL;
      du[N + 8] = VN;
    }
    __syncwarp(LFULL);
  }
};

// the scalar core of the daily callback (lane 0): out of line, so that its registers do not burden the stepping loop
static __device__ __noinline__ void wide_cb(const ParamView& pv, const double* __restrict__ frow, double t, double psim0,
                                            const double* psiti, const double* kk, DayState& ds, double* TRANI, double* out4,
                                            double* TRANIst, double& RWUst, int& cbst) {
  cb_core<1>(pv, frow, t, psim0, psiti, kk, ds, TRANI, out4[0], out4[1], out4[2], out4[3], TRANIst, RWUst, cbst);
}

// The whole integration of member m by one warp.  `smem`: the member's block of LWFB_WIDE_DOUBLES_PER_MEMBER(N) doubles.
__device__ __forceinline__ void wide_body(const DevData& d, const int lane, const int64_t m, double* smem) {
  Wide mb;
  mb.load(d, m, lane, smem);
  const int N = mb.N, NE = N + 9, L = d.L;
  const bool iso = d.iso != 0;
  const SIdx ix(N, d.iso);
  const double* __restrict__ forc = d.forc + (int64_t)__ldg(d.site + m) * d.n_rows * LWFB_F;
  auto U0 = [&](int j) { return __ldg(d.u0 + m * L + j); };
  double *y = mb.A(WA_X), *yprev = mb.A(WA_Y);
  double *dS18 = mb.A(WA_DS18), *dS2H = mb.A(WA_DS2H), *TRANIst = mb.A(WA_TRANIST);
  auto K = [&](int s) { return mb.A(WA_K0 + s); };
  const bool l0 = lane == 0;

  // member-level state that only lane 0 keeps (the other lanes never read theirs)
  DayState ds;
  double RWUst = 0, XYmm = 0, dRWU[2] = {0, 0}, dISN[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  // member-level state every lane keeps (computed from identical inputs)
  double dG[2] = {0, 0}, dXY[2] = {0, 0};
  for (int i = lane; i < N; i += 32) {
    y[i] = U0(ix.SWATI + i); TRANIst[i] = U0(ix.TRANI + i);
    if (d.ntfli0) mb.A(WA_NTFLI)[i] = __ldg(d.ntfli0 + m * N + i);     // warm du_NTFLI cache
    dS18[i] = iso ? U0(ix.SWATI + N + i) : 0.0; dS2H[i] = iso ? U0(ix.SWATI + 2 * N + i) : 0.0;
  }
  for (int a = 0; a < LWFB_NACCUM; a++) ds.acc[a] = U0(ix.ACC + a);
  if (l0) { y[N] = U0(ix.GWAT); for (int q = 0; q < 8; q++) y[N + 1 + q] = ds.acc[kAccPos[q]]; }
  ds.INTS = U0(ix.INTS); ds.INTR = U0(ix.INTR); ds.SNOW = U0(ix.SNOW); ds.CC = U0(ix.CC); ds.SNOWLQ = U0(ix.SNOWLQ);
  RWUst = U0(ix.RWU); XYmm = U0(ix.XYLEM);
  if (iso) {
    for (int s = 0; s < 2; s++) {
      dG[s] = U0(ix.GWAT + 1 + s); dRWU[s] = U0(ix.RWU + 1 + s); dXY[s] = U0(ix.XYLEM + 1 + s);
      dISN[0][s] = U0(ix.INTS + 1 + s); dISN[1][s] = U0(ix.INTR + 1 + s); dISN[2][s] = U0(ix.SNOW + 1 + s);
    }
  }
  __syncwarp(LFULL);

  // the full reference-order state record (src/func_DiffEq_definition_u0.jl:42-73): per-layer entries by all lanes,
  // member-level entries by lane 0
  auto write_record = [&](double* __restrict__ dst) {
    for (int i = lane; i < N; i += 32) {
      dst[ix.SWATI + i] = y[i];
      dst[ix.TRANI + i] = TRANIst[i];
      if (iso) {
        dst[ix.SWATI + N + i] = dS18[i]; dst[ix.SWATI + 2 * N + i] = dS2H[i];
        dst[ix.TRANI + N + i] = U0(ix.TRANI + N + i); dst[ix.TRANI + 2 * N + i] = U0(ix.TRANI + 2 * N + i);   // never updated by the reference
      }
    }
    for (int j = lane; j < 3 * N; j += 32) dst[ix.AUX + j] = U0(ix.AUX + j);      // aux is never written after init
    if (l0) {
      dst[ix.GWAT] = y[N]; dst[ix.INTS] = ds.INTS; dst[ix.INTR] = ds.INTR; dst[ix.SNOW] = ds.SNOW;
      dst[ix.RWU] = RWUst; dst[ix.XYLEM] = XYmm; dst[ix.CC] = ds.CC; dst[ix.SNOWLQ] = ds.SNOWLQ;
      if (iso) {
        for (int s = 0; s < 2; s++) {
          dst[ix.GWAT + 1 + s] = dG[s]; dst[ix.RWU + 1 + s] = dRWU[s]; dst[ix.XYLEM + 1 + s] = dXY[s];
          dst[ix.INTS + 1 + s] = dISN[0][s]; dst[ix.INTR + 1 + s] = dISN[1][s]; dst[ix.SNOW + 1 + s] = dISN[2][s];
        }
      }
      for (int a = 0; a < LWFB_NACCUM; a++) dst[ix.ACC + a] = ds.acc[a];
      for (int q = 0; q < 8; q++) dst[ix.ACC + kAccPos[q]] = y[N + 1 + q];
    }
  };
  // goodness-of-fit sums at the observation days (lwfb_set_observations), from what daily record r holds (lane 0)
  auto score_rec = [&](int r, bool reached) {
    if (d.n_obs_layers == 0 || !l0) return;
    const int ko = __ldg(d.obs_of_rec + r);
    if (ko < 0) return;
    double* a = d.sacc + m * (2 * LWFB_SA);
    for (int l = 0; l < d.n_obs_layers; l++) {
      const int il = __ldg(d.obs_layer + l);
      const double sw = reached ? (r == 0 ? U0(ix.SWATI + il) : y[il]) : NAN;
      if (d.obs_theta)
        score_add(a, reached ? score_theta(mb.A(WA_THSAT)[il], mb.A(WA_THR)[il], mb.A(WA_SWATMAX)[il], sw) : NAN,
                  __ldg(d.obs_theta + (int64_t)ko * d.n_obs_layers + l));
      if (d.obs_d18o && iso)
        score_add(a + LWFB_SA, reached ? (r == 0 ? U0(ix.SWATI + N + il) : dS18[il]) : NAN,
                  __ldg(d.obs_d18o + (int64_t)ko * d.n_obs_layers + l));
    }
  };
  // delta of precipitation at time t: the value of forcing row r holds for t in (r - 0.5, r + 0.5]
  auto prec_delta = [&](double t, double* dp) {
    int r = (int)ceil(t - 0.5);
    r = r < 0 ? 0 : (r > d.n_rows - 1 ? d.n_rows - 1 : r);
    dp[0] = __ldg(forc + (int64_t)r * LWFB_F + LWFB_F_D18O);
    dp[1] = __ldg(forc + (int64_t)r * LWFB_F + LWFB_F_D2H);
  };
  int status = LWFB_ST_OK;
  // daily callback at time t (src/func_DiffEq_definition_cb.jl:60-302, :304-530): soil state of the (clamped) layers by
  // all lanes, the scalar core by lane 0, its results that every lane needs broadcast
  auto daily_callback = [&](double t, int day) {
    double psim0 = 0;
    for (int i = lane; i < N; i += 32) {
      double W, psim, th, kki;
      mb.soilvar1(i, y, W, psim, th, kki);
      mb.A(WA_KK)[i] = kki;
      mb.A(WA_PSITI)[i] = psim + mb.A(WA_PSIG)[i];
      if (i == 0) psim0 = psim;
    }
    __syncwarp(LFULL);
    double o4[4] = {0, 0, 0, 0};               // RNET, SMLT, SLVP, TA
    int cbst = LWFB_ST_OK;
    if (l0) {
      wide_cb(mb.pv, forc + (int64_t)day * LWFB_F, t, psim0, mb.A(WA_PSITI), mb.A(WA_KK), ds, mb.A(WA_TRANI), o4, TRANIst, RWUst, cbst);
      for (int q8 = 0; q8 < 8; q8++) y[N + 1 + q8] = 0.0;
      if (iso) {
        double dp[2];
        prec_delta(t, dp);
        for (int c = 0; c < 3; c++) { dISN[c][0] = dp[0]; dISN[c][1] = dp[1]; }
      }
    }
    __syncwarp(LFULL);
    mb.RS = __shfl_sync(LFULL, o4[0] + o4[1], 0);
    mb.SLVP = __shfl_sync(LFULL, o4[2], 0);
    mb.TA = __shfl_sync(LFULL, o4[3], 0);
    cbst = __shfl_sync(LFULL, cbst, 0);
    if (cbst != LWFB_ST_OK) status = cbst;
  };
  // LWFBrook90R_updateIsotopes_GWAT_SWAT_AdvecDiff! src/func_DiffEq_definition_cb.jl:635-1250 after an accepted step
  auto ade = [&](double dt, const double* dprec) {
    const double* S1 = y;
    const double G1 = *(const volatile double*)(y + N);
    double *th1 = mb.A(WA_T0), *tauth = mb.A(WA_T0 + 1), *aq = mb.A(WA_T0 + 2), *C = mb.A(WA_T0 + 3), *Dd = mb.A(WA_T0 + 4),
           *Ctr = mb.A(WA_T0 + 5), *ct = mb.A(WA_T0 + 6);
    const double *TRANI = mb.A(WA_TRANI), *VRFLI = mb.A(WA_VRFLI), *INFLI = mb.A(WA_INFLI), *DSFLI = mb.A(WA_DSFLI),
                 *NTFLI = mb.A(WA_NTFLI);
    const double Tk = mb.TA + 273.15;
    const double Lam = mb.pv.S(LWFB_S_DISPERSIVITY), VXYLEM = mb.pv.S(LWFB_S_VXYLEM);
    const double e = m_exp(-535400 / (Tk * Tk) + 1393.3 / Tk + 2.1876);
    for (int i = lane; i < N; i += 32) {
      const double THSAT = mb.A(WA_THSAT)[i], THR = mb.A(WA_THR)[i];
      const double w = jmin(1.0, div_c(div_c(THSAT * S1[i], mb.A(WA_SWATMAX)[i]) - THR, THSAT - THR));
      th1[i] = w * (THSAT - THR) + THR;
      tauth[i] = m_pow(th1[i], 7.0 / 3) * LWFB_RCP(THSAT * THSAT);
      aq[i] = Lam * fabs(VRFLI[i] * LWFB_RCP(1000.0));
    }
    __syncwarp(LFULL);
    const double sumT = mb.ordered_sum(TRANI, N);
    for (int s = 0; s < 2; s++) {
      const double R = s == 0 ? R18 : R2H;
      const double D0 = (s == 0 ? 0.96691 : 0.98331) * 1e-9 * e * 3600 * 24;
      double* dk = s == 0 ? dS18 : dS2H;
      const double Cxy = d_to_x(dXY[s], R);
      const double Cinf = d_to_x(dprec[s], R);
      for (int i = lane; i < N; i += 32) {
        Dd[i] = D0 * tauth[i] * th1[i] + aq[i];
        C[i] = d_to_x(dk[i], R);
        Ctr[i] = (TRANI[i] < 0) ? Cxy : C[i];
        ct[i] = Ctr[i] * TRANI[i];
      }
      __syncwarp(LFULL);
      const double Cslvp = *(const volatile double*)(C + 0);
      for (int i = lane; i < N; i += 32) {
        const double diff_upp = (i == 0) ? 0.0 : ((Dd[i - 1] + Dd[i]) / 2) * ((C[i] - C[i - 1]) * LWFB_RCP(mb.pv.Lp(LWFB_L_DZ, i - 1)));
        const double diff_low = (i == N - 1) ? 0.0 : ((Dd[i] + Dd[i + 1]) / 2) * ((C[i + 1] - C[i]) * LWFB_RCP(mb.pv.Lp(LWFB_L_DZ, i)));
        const double qC_upp = (i == 0) ? 0.0 : VRFLI[i - 1] * C[i - 1];
        const double qC_low = VRFLI[i] * C[i];
        const double slv = (i == 0) ? Cslvp : 0.0;
        double inner = -diff_upp * 1000 + diff_low * 1000 + qC_upp;
        inner = inner - qC_low;
        inner = inner + INFLI[i] * Cinf;
        inner = inner - DSFLI[i] * C[i];
        inner = inner - mb.SLVP * slv;
        inner = inner - TRANI[i] * Ctr[i];
        const double dC = -C[i] * (1 / S1[i]) * NTFLI[i] + 1 / S1[i] * inner;
        dk[i] += dt * dxdt_to_dddt(dC, C[i], R);
      }
      double ddG;
      const double VN = *(const volatile double*)(VRFLI + N - 1);
      if ((G1 == 0) && (VN == 0)) ddG = NAN;
      else if ((G1 == 0) && (VN > 0)) ddG = 0;
      else {
        const double CG = d_to_x(dG[s], R);
        const double dCG = VN / G1 * (*(const volatile double*)(C + N - 1) - CG);
        ddG = dxdt_to_dddt(dCG, CG, R);
      }
      dG[s] += dt * ddG;
      const double sw = mb.ordered_sum(ct, N);
      const double dRW = x_to_d(sw / sumT, R);
      dRWU[s] = dRW;
      const double ddxy = (dRW != dRW) ? 0.0 : sumT / VXYLEM * (dRW - dXY[s]);
      dXY[s] += dt * ddxy;
      __syncwarp(LFULL);             // C, Dd, Ctr, ct are rewritten for the second species
    }
  };

  const double qmin = 0.2, qmax = 10.0, gamma = 0.9, beta1 = 0.14, beta2 = 0.08, qoldinit = 1e-4;
  const double t0 = (double)d.t0_day, tend = (double)(d.t0_day + d.n_days);
  double t = t0, tprev = t0, dt = d.dt_init, dtpropose = d.dt_init, qold = qoldinit, q11 = 1.0;
  const double dtmin_opt = fmax(eps_of(tend), eps_of(t0));
  long long naccept = 0, nreject = 0, iter = 0;
  bool accept = false, reeval = false;

  const bool probe = d.probe_out != nullptr;         // lwfb_eval_rhs: f(u, t) with the given frozen fluxes, nothing else
  if (probe) {
    const double* fr = d.probe_frozen + m * (N + 3);
    mb.RS = __ldg(fr) + __ldg(fr + 1); mb.SLVP = __ldg(fr + 2);
    for (int i = lane; i < N; i += 32) {
      mb.A(WA_TRANI)[i] = __ldg(fr + 3 + i); mb.A(WA_NTFLI)[i] = __ldg(d.probe_ntfli + m * N + i);
      yprev[i] = U0(ix.SWATI + i);
    }
    if (l0) yprev[N] = U0(ix.GWAT);
    __syncwarp(LFULL);
    mb.rhs(yprev, d.probe_t, K(0));
    for (int j = lane; j < NE; j += 32) d.probe_out[m * NE + j] = K(0)[j];
    return;
  }

  // ---- init: save u0, initial_affect of the daily callbacks, FSAL
  double* const rec_m = d.rec ? d.rec + m * (int64_t)(d.n_days + 1) * L : nullptr;   // [n_days+1][L] of this member
  if (d.rec) write_record(rec_m);
  score_rec(0, true);
  daily_callback(t0, d.t0_day);
  __syncwarp(LFULL);
  for (int j = lane; j < NE; j += 32) yprev[j] = y[j];
  __syncwarp(LFULL);
  mb.rhs(yprev, t, K(0));
  double tstop = fmin(t0 + 1.0, tend);
  int rec = 1;

  while (status == LWFB_ST_OK && t < tend) {
    // loopheader!
    if (iter > 0) {
      if (accept) {
        for (int j = lane; j < NE; j += 32) yprev[j] = y[j];
        dt = dtpropose;
        __syncwarp(LFULL);
        if (reeval) mb.rhs(y, t, K(0));
        else { for (int j = lane; j < NE; j += 32) K(0)[j] = K(6)[j]; __syncwarp(LFULL); }
      } else {
        dt = dt / jmin(1 / qmin, q11 / gamma);
      }
    }
    iter++;
    if (iter > d.maxiters) { status = LWFB_ST_MAXITERS; break; }
    {
      bool bad = false;
      for (int i = lane; i < N; i += 32) bad |= (y[i] != y[i]);
      if (__any_sync(LFULL, bad)) { status = LWFB_ST_UNSTABLE; break; }
    }
    dt = dt < d.dtmax ? dt : d.dtmax;
    { const double dtmin_t = fmax(eps_of(t), dtmin_opt); if (dt < dtmin_t) dt = dtmin_t; }
    if (dt > tstop - t) dt = tstop - t;
    if (!(fabs(dt) > dtmin_opt)) { status = LWFB_ST_DTMIN; break; }
    // perform_step! Tsit5; only SWATI and GWAT feed the RHS, the accumulators ride along
    for (int s = 1; s <= 5; s++) {
      for (int j = lane; j <= N; j += 32) {
        if (s == 1) y[j] = yprev[j] + (dt * a21) * K(0)[j];
        else {
          double acc = kA[s][0] * K(0)[j];
          for (int l = 1; l < s; l++) acc = acc + kA[s][l] * K(l)[j];
          y[j] = yprev[j] + dt * acc;
        }
      }
      __syncwarp(LFULL);
      mb.rhs(y, (s < 5) ? t + kC[s] * dt : t + dt, K(s));
    }
    for (int j = lane; j < NE; j += 32)
      y[j] = yprev[j] + dt * (a71 * K(0)[j] + a72 * K(1)[j] + a73 * K(2)[j] + a74 * K(3)[j] + a75 * K(4)[j] + a76 * K(5)[j]);
    __syncwarp(LFULL);
    mb.rhs(y, t + dt, K(6));         // clamps the new SWATI in place
    // residuals + norm_to_use (src/func_DiffEq_definition_ode.jl:20-37): RMS over 3N+39 entries, summed in index order
    double* r2 = mb.A(WA_T0);
    for (int j = lane; j < NE; j += 32) {
      const double ut = dt * (bt1 * K(0)[j] + bt2 * K(1)[j] + bt3 * K(2)[j] + bt4 * K(3)[j] + bt5 * K(4)[j] + bt6 * K(5)[j] + bt7 * K(6)[j]);
      const double r = ut / (d.abstol + jmax(fabs(yprev[j]), fabs(y[j])) * d.reltol);
      r2[j] = r * r;
    }
    __syncwarp(LFULL);
    const double ss = mb.ordered_sum(r2, NE);
    __syncwarp(LFULL);
    const double EEst = sqrt(ss / (double)(3 * N + 39));
    // loopfooter!: PI controller
    reeval = false;
    const double ttmp = t + dt;
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
      const double dtnew = dt / q;
      tprev = t;
      t = (fabs(ttmp - tstop) < 100 * eps_of(fmax(t, tstop))) ? tstop : ttmp;
      dtpropose = d.dtmax < dtnew ? d.dtmax : dtnew;
      { const double dtmin_t = fmax(eps_of(t), dtmin_opt); if (dtpropose < dtmin_t) dtpropose = dtmin_t; }
      double dp[2] = {0, 0};
      if (iso) prec_delta(t, dp);
      if (t == tstop) {
        // the record is saved before the callbacks of this instant (and before this step's ADE update)
        if (d.rec) write_record(rec_m + (int64_t)rec * L);
        score_rec(rec, true);
        rec++;
        if (t < tend) {
          // LWFBrook90R_check_balance_errors! src/func_DiffEq_definition_cb.jl:1255-1314
          const double Ssum = mb.ordered_sum(y, N);
          if (l0) {
            double* a = ds.acc;
            const double Wt = ds.INTR + ds.INTS + ds.SNOW + Ssum + y[N];
            const double qin = 0 + (y[N + 4] - y[N + 5]);
            const double qout = y[N + 8] + y[N + 6] + a[LWFB_A_CUM_D_TRAN] + a[LWFB_A_CUM_D_SLVP];
            const double qin_t = a[LWFB_A_CUM_D_PREC];
            const double qout_t = a[LWFB_A_CUM_D_EVAP] + y[N + 1] + y[N + 2];
            a[LWFB_A_BALERD_SWAT] = (qin - qout) - (Ssum - a[LWFB_A_STORAGESWAT]);
            a[LWFB_A_BALERD_TOTAL] = (qin_t - qout_t) - (Wt - a[LWFB_A_STORAGEWATER]);
            a[LWFB_A_STORAGESWAT] = Ssum;
            a[LWFB_A_STORAGEWATER] = Wt;
          }
          if (iso) ade(t - tprev, dp);
          __syncwarp(LFULL);
          daily_callback(t, (int)floor(t + 0.5));
          reeval = true;
          tstop = fmin(tstop + 1.0, tend);
        }
      } else {
        if (iso) ade(t - tprev, dp);
      }
      __syncwarp(LFULL);
    } else {
      nreject++;
    }
  }
  __syncwarp(LFULL);
  for (int r = rec; r <= d.n_days; r++) score_rec(r, false);
  if (d.rec) {
    for (; rec <= d.n_days; rec++)
      for (int j = lane; j < L; j += 32) rec_m[(int64_t)rec * L + j] = NAN;
  }
  write_record(d.fin + m * L);
  if (d.ntfli_out) for (int i = lane; i < N; i += 32) d.ntfli_out[m * N + i] = mb.A(WA_NTFLI)[i];
  if (l0) {
    d.status[m] = status;
    d.counters[m * LWFB_NCOUNTERS + LWFB_C_NACCEPT] = naccept;
    d.counters[m * LWFB_NCOUNTERS + LWFB_C_NREJECT] = nreject;
    d.counters[m * LWFB_NCOUNTERS + LWFB_C_NF] = mb.nf;
  }
}

}  // namespace lwfb
