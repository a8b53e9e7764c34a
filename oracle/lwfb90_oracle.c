/*
 * lwfb90_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, one-member-at-a-time restatement of the algorithm the reference (fabern/LWFBrook90.jl
 * v0.9.9) executes inside simulate!() -> solve().  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may build, load or call this file.  The product
 * path (lwfbrook90.jl_b200/csrc) never links against it.
 *
 * Parity status: the reference cannot run in the build container (no Julia).  The oracle is
 * pinned against the reference's own golden CSVs (examples/DAV2020-full_u_sol_reference*.csv,
 * examples/INFEXP1*_fluxes_referencedf.csv, examples/DAV2020-full-modified_fluxes_referencedf.csv)
 * and inline tables (test/01-unit-tests.jl:907-945, :951-1014) through committed fixtures in
 * tests/golden/.  The time integrator itself (OrdinaryDiffEq.jl Tsit5 + PI controller,
 * DiffEqCallbacks.jl PeriodicCallback/FunctionCallingCallback, DiffEqBase fastpow) is a
 * third-party dependency that is absent from /root/reference and unpinned (Project.toml:34-56,
 * no Manifest); its published algorithm is restated below (section "SOLVER").
 *
 * Every function cites the reference file:line it follows.  Paths are relative to the
 * reference root.
 */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/lwfb90.h"

/* Build mode B (-DLWFB_ORACLE_ELEM): identical code, but the elementary functions are the portable
 * routines the CUDA kernels use (include/lwfb90_elem.h), so that GPU results can be compared
 * BIT-EXACTLY.  Mode A (default) uses the host libm and is the one pinned against the goldens. */
#ifdef LWFB_ORACLE_ELEM
#include "../include/lwfb90_elem.h"
#define exp lwfb_exp
#define log lwfb_log
#define pow lwfb_pow
#define sin lwfb_sin
#define cos lwfb_cos
#define tan lwfb_tan
#define asin lwfb_asin
#define acos lwfb_acos
#endif

/* Mode B also forms the quotients by run constants the way the kernels do (csrc/lwfb90_device.cuh, div_c / LWFB_RCP):
 * DIVX in the water path = product with the IEEE reciprocal plus one exact correction step (the correctly rounded
 * quotient for finite operands), DIVC in the isotope step = plain product with the reciprocal.  Mode A divides, like
 * the reference. */
#if defined(LWFB_ORACLE_ELEM) && !defined(LWFB_ORACLE_NORCP)
static double divx_(double x, double c) {
  const double rc = 1.0 / c;
  const double q = x * rc;
  const double r = fma(-c, q, x);
  const double e = fma(r, rc, q);
  return (fabs(x) < INFINITY) ? e : q;
}
#define DIVX(x, c) divx_((x), (c))
#define DIVC(x, c) ((x) * (1.0 / (c)))
#else
#define DIVX(x, c) ((x) / (c))
#define DIVC(x, c) ((x) / (c))
#endif

#define MAXN 512

/* ---- src/module_CONSTANTS.jl:5-23 ------------------------------------------------------ */
static const double p_WTOMJ = 0.0864, p_ETOM = 0.4085, p_CPRHO = 1240., p_GAMMA = 0.067,
                    p_CVLQ = 0.00418, p_CVICE = 0.00192, p_LF = 0.335, p_LS = 2.824,
                    p_RHOWG = 0.00981, p_SIGMA = 5.67E-08, p_SC = 1367., p_K = 0.4,
                    p_PI = 3.1416, p_DTIMIN = 1.e-9;

/* Julia semantics: min/max propagate NaN */
static double jmin(double a, double b) { if (isnan(a) || isnan(b)) return NAN; return a < b ? a : b; }
static double jmax(double a, double b) { if (isnan(a) || isnan(b)) return NAN; return a > b ? a : b; }

typedef struct {
  int N, iso;
  const double* scal;  /* [LWFB_PS]     */
  const double* lay;   /* [LWFB_PL][N]  */
  const double* forc;  /* [rows][LWFB_F]*/
  int n_rows;
  /* mutable caches of the reference's `p` (src/func_DiffEq_definition_p.jl:503-581) */
  double RNET, SMLT, SLVP, TA;
  double TRANI[MAXN];
  double NTFLI[MAXN], VRFLI[MAXN], INFLI[MAXN], BYFLI[MAXN], DSFLI[MAXN];
  double GWFL, SEEP, SRFL, SLFL;
  int status;
  int64_t nf;
} model_t;

#define SC(m, k) ((m)->scal[LWFB_S_##k])
#define LY(m, k, i) ((m)->lay[(size_t)LWFB_L_##k * (m)->N + (i)])
#define FC(m, d, k) ((m)->forc[(size_t)(d) * LWFB_F + LWFB_F_##k])

/* =========================================================================================
 * KPT — src/module_KPT.jl
 * ========================================================================================= */

/* SWCHEK!  src/module_KPT.jl:761-790 */
static void swchek(const model_t* m, double* swati) {
  for (int i = 0; i < m->N; i++) {
    if (swati[i] <= 0.0) swati[i] = 0.0;
    else if (swati[i] > LY(m, SWATMAX, i)) swati[i] = LY(m, SWATMAX, i);
  }
}

/* derive_auxiliary_SOILVAR!  src/module_KPT.jl:445-464; FPSIM_MvG! :596-605; FK_MvG! :502-509;
 * FTheta_MvG! :692-696 */
static void soilvar(const model_t* m, const double* swati, double* W, double* psim, double* psiti,
                    double* theta, double* kk) {
  for (int i = 0; i < m->N; i++) {
    double ths = LY(m, THSAT, i), thr = LY(m, THR, i);
    double w = jmin(1.0, DIVX(DIVX(ths * swati[i], LY(m, SWATMAX, i)) - thr, ths - thr));
    W[i] = w;
    double a = jmax(w, 1.e-6);
    double n = LY(m, NPAR, i), mm = LY(m, MPAR, i);
    psim[i] = 9.81 * (-1.0 / LY(m, ALPHA, i)) * pow(pow(a, -1.0 / mm) - 1.0, 1.0 / n);
    theta[i] = w * (ths - thr) + thr;
    double inner = 1.0 - pow(1.0 - pow(a, n / (n - 1.0)), 1.0 - 1.0 / n);
    kk[i] = LY(m, KSAT, i) * pow(a, LY(m, TORT, i)) * (inner * inner);
    psiti[i] = psim[i] + LY(m, PSIG, i);
  }
}

/* =========================================================================================
 * WAT — src/module_WAT.jl
 * ========================================================================================= */

/* KKMEAN :238-252 and VERT :206-229 */
static double vert(double KK_i, double KK_ip, double KSAT_i, double KSAT_ip, double T_i, double T_ip,
                   double PSIT_i, double PSIT_ip, double ST_i, double ST_ip) {
  double kkmean = exp(DIVX(T_ip * log(KK_i) + T_i * log(KK_ip), T_i + T_ip));
  kkmean = jmin(jmin(kkmean, KSAT_i), KSAT_ip);
  double grad = DIVX(PSIT_i - PSIT_ip, (T_i + T_ip) / 2.0);
  return DIVX(grad * kkmean, p_RHOWG) * (1.0 - (ST_i + ST_ip) / 2.0);
}

/* INFLOW!  src/module_WAT.jl:313-457 */
static void inflow(const model_t* m, double* VV, double* INFLI, double* BYFLI, double DTI,
                   const double* BYFRAC, double SLFL, const double* DSFLI, const double* TRANI,
                   double SLVP, const double* swati, const double* VRFLI_prior) {
  int N = m->N;
  for (int i = N - 1; i >= 0; i--) {
    BYFLI[i] = SLFL * LY(m, INFRAC, i) * BYFRAC[i];
    INFLI[i] = SLFL * LY(m, INFRAC, i) * (1.0 - BYFRAC[i]);
  }
  VV[N - 1] = VRFLI_prior[N - 1];
  for (int i = N - 1; i >= 0; i--) {
    double MAXIN = (LY(m, SWATMAX, i) - swati[i]) / DTI + VV[i] + DSFLI[i] + TRANI[i];
    if (i == 0) {
      MAXIN = MAXIN + SLVP;
      if (INFLI[0] > MAXIN) {
        double too = INFLI[0] - MAXIN;
        INFLI[0] = INFLI[0] - too;
        BYFLI[0] = BYFLI[0] + too;
      }
    } else {
      if (VRFLI_prior[i - 1] + INFLI[i] > MAXIN) {
        if (BYFRAC[i] > 0) {
          if (VRFLI_prior[i - 1] < MAXIN) {
            VV[i - 1] = VRFLI_prior[i - 1];
            double too = INFLI[i] + VRFLI_prior[i - 1] - MAXIN;
            INFLI[i] = INFLI[i] - too;
            BYFLI[i] = BYFLI[i] + too;
          } else {
            VV[i - 1] = MAXIN;
            BYFLI[i] = BYFLI[i] + INFLI[i];
            INFLI[i] = 0;
          }
        } else {
          VV[i - 1] = MAXIN - INFLI[i];
        }
      } else {
        VV[i - 1] = VRFLI_prior[i - 1];
      }
    }
  }
}

/* ITER (Mualem-van Genuchten branch)  src/module_WAT.jl:459-571 */
static double iter_mvg(const model_t* m, double DTI, const double* NTFLI_prev, const double* theta) {
  double DTINEW = DTI;
  double DSWMAX = SC(m, DSWMAX);
  for (int i = 0; i < m->N; i++) {
    double T = LY(m, THICK, i);
    DTINEW = jmin(DTINEW, 0.01 * DSWMAX * T * LY(m, THSAT, i) * (1 - LY(m, STONEF, i)) /
                              jmax(0.000001, fabs(NTFLI_prev[i])));
    double avail = (LY(m, THR, i) - theta[i]) * T;
    if (NTFLI_prev[i] < 0) {
      DTINEW = jmin(DTINEW, DIVX(avail / NTFLI_prev[i], 1.30));
      if (DTINEW < p_DTIMIN) DTINEW = p_DTIMIN;
    }
  }
  return DTINEW;
}

/* GWATER  src/module_WAT.jl:578-596 */
static void gwater(double GWAT, double GSC, double GSP, double VRFLN, double* GWFL, double* SEEP) {
  if (GSC < 1.0e-8) {
    *SEEP = GSP * VRFLN;
    *GWFL = VRFLN - *SEEP;
  } else {
    *SEEP = GWAT * GSC * GSP;
    *GWFL = GWAT * GSC * (1. - GSP);
    if (GWAT - (*GWFL + *SEEP) * 1 < 0) {
      *SEEP = GWAT / 1 * GSP;
      *GWFL = GWAT / 1 * (1. - GSP);
    }
  }
}

/* =========================================================================================
 * RHS  f_LWFBrook90!  src/func_DiffEq_definition_f.jl:7-243 with MSBITERATE!
 * src/func_MSB_functions.jl:565-697.  du = (dSWATI[N], dGWAT, flow, seep, srfl, slfl, byfl,
 * dsfl, gwfl, vrfln).  swati is clamped IN PLACE.
 * ========================================================================================= */
static void rhs(model_t* m, double* swati, double gwat, double t, double* du) {
  int N = m->N;
  double W[MAXN], psim[MAXN], psiti[MAXN], theta[MAXN], kk[MAXN], BYFRAC[MAXN], V0[MAXN];
  m->nf++;
  swchek(m, swati);                                              /* f.jl:61 */
  soilvar(m, swati, W, psim, psiti, theta, kk);                  /* f.jl:66 */
  /* bypass fraction f.jl:78-89, BYFLFR! module_WAT.jl:133-142 */
  if (SC(m, BYPAR) == 1.0) {
    for (int i = 0; i < N; i++) {
      double wf = LY(m, WETF, i);
      BYFRAC[i] = jmin(1.0, pow(SC(m, QFFC), 1 - (1 / SC(m, QFPAR)) * (W[i] - wf) / (1 - wf)));
      if (W[i] > 0.99) BYFRAC[i] = 1;
    }
  } else {
    for (int i = 0; i < N; i++) BYFRAC[i] = 0;
  }
  /* MSBITERATE! MSB:594-602 */
  double SAFRAC = 0.;
  int QLAYER = (int)SC(m, QLAYER);
  if (QLAYER > 0) { /* SRFLFR module_WAT.jl:144-150 */
    double s = 0;
    for (int i = 0; i < QLAYER; i++) s += swati[i];
    SAFRAC = jmin(1.0, pow(SC(m, QFFC), 1 - (1 / SC(m, QFPAR)) * (s - SC(m, SWATQF)) / (SC(m, SWATQX) - SC(m, SWATQF))));
  }
  m->SRFL = jmin(1., (SC(m, IMPERV) + SAFRAC)) * (m->RNET + m->SMLT);
  m->SLFL = m->RNET + m->SMLT - m->SRFL;
  /* DSLOP! MSB:607-614, module_WAT.jl:178-198 (DSLOPE used as radians) */
  if (SC(m, LENGTH_SLOPE) == 0 || SC(m, DSLOPE) == 0) {
    for (int i = 0; i < N; i++) m->DSFLI[i] = 0;
  } else {
    double ds = SC(m, DSLOPE), len = SC(m, LENGTH_SLOPE);
    for (int i = 0; i < N; i++) {
      double GRAD = cos(ds) * (psim[i] * 2 / (1000 * len)) + sin(ds) * p_RHOWG;
      double ARATIO = LY(m, THICK, i) * (1 - LY(m, STONEF, i)) * cos(ds) / (1000 * len);
      double d = kk[i] * ARATIO * GRAD / p_RHOWG;
      m->DSFLI[i] = d < 0 ? 0 : d;
    }
  }
  /* first approximation of vertical flow MSB:617-641 */
  for (int i = N - 1; i >= 0; i--) {
    if (i < N - 1) {
      if (fabs(psiti[i] - psiti[i + 1]) < SC(m, DPSIMAX)) V0[i] = 0;
      else
        V0[i] = vert(kk[i], kk[i + 1], LY(m, KSAT, i), LY(m, KSAT, i + 1), LY(m, THICK, i), LY(m, THICK, i + 1),
                     psiti[i], psiti[i + 1], LY(m, STONEF, i), LY(m, STONEF, i + 1));
    } else {
      if (SC(m, DRAIN) > 0.00001) V0[N - 1] = SC(m, DRAIN) * kk[N - 1] * (1 - LY(m, STONEF, N - 1));
      else V0[N - 1] = 0;
    }
  }
  /* MSB:644-646; p_DTP = 1 */
  double DTRI = 1.0 - fmod(t, 1.0);
  double DTI = jmin(DTRI, SC(m, DTIMAX));
  inflow(m, m->VRFLI, m->INFLI, m->BYFLI, DTI, BYFRAC, m->SLFL, m->DSFLI, m->TRANI, m->SLVP, swati, V0);
  /* ITER consumes the du_NTFLI of the PREVIOUS call (MSB:668-670).
   * FDPSIDWF! (MSB:666) is dead code for MvG. */
  double DTINEW = iter_mvg(m, DTI, m->NTFLI, theta);
  if (DTINEW < DTI) { /* false when DTINEW is NaN */
    DTI = DTINEW;
    inflow(m, m->VRFLI, m->INFLI, m->BYFLI, DTI, BYFRAC, m->SLFL, m->DSFLI, m->TRANI, m->SLVP, swati, V0);
  }
  for (int i = N - 1; i >= 0; i--) { /* MSB:685-693 */
    if (i == 0) m->NTFLI[i] = m->INFLI[i] - m->VRFLI[i] - m->DSFLI[i] - m->TRANI[i] - m->SLVP;
    else m->NTFLI[i] = m->VRFLI[i - 1] + m->INFLI[i] - m->VRFLI[i] - m->DSFLI[i] - m->TRANI[i];
  }
  gwater(gwat, SC(m, GSC), SC(m, GSP), m->VRFLI[N - 1], &m->GWFL, &m->SEEP); /* f.jl:117 */
  double sby = 0, sds = 0;
  for (int i = 0; i < N; i++) sby += m->BYFLI[i];
  for (int i = 0; i < N; i++) sds += m->DSFLI[i];
  for (int i = 0; i < N; i++) du[i] = m->NTFLI[i];               /* f.jl:162 */
  du[N + 0] = m->VRFLI[N - 1] - m->GWFL - m->SEEP;               /* f.jl:143 */
  du[N + 1] = m->SRFL + sby + sds + m->GWFL;                     /* flow  f.jl:207-210 */
  du[N + 2] = m->SEEP;
  du[N + 3] = m->SRFL;
  du[N + 4] = m->SLFL;
  du[N + 5] = sby;
  du[N + 6] = sds;
  du[N + 7] = m->GWFL;
  du[N + 8] = m->VRFLI[N - 1];
}

/* =========================================================================================
 * SUN — src/module_SUN.jl
 * ========================================================================================= */
static double hafday(double LAT, double DEC) { /* :194-207 */
  if (fabs(LAT) >= p_PI / 2) LAT = (LAT > 0 ? 1.0 : (LAT < 0 ? -1.0 : 0.0)) * (p_PI / 2 - 0.01);
  double ARG = -tan(DEC) * tan(LAT);
  if (ARG >= 1) return 0;
  else if (ARG <= -1) return p_PI;
  return acos(ARG);
}
static double func3(double DEC, double L2, double L1, double a3, double a2) { /* :189-192 */
  return (1.0 / (2.0 * p_PI)) *
         (sin(DEC) * sin(L1) * (a3 - a2) + cos(DEC) * cos(L1) * (sin(a3 + L2) - sin(a2 + L2)));
}
static void sunds(double LAT, double SLOPE, double DOY, double L1, double L2, double* DAYLEN,
                  double* I0HDAY, double* SLFDAY) { /* :140-187 */
  double d = 1.0 - 0.0167 * cos(0.0172 * (DOY - 3));
  double SCD = p_SC / (d * d);
  double DEC = asin(.39785 * sin(4.868961 + 0.017203 * DOY + 0.033446 * sin(6.224111 + 0.017202 * DOY)));
  double temp = hafday(LAT, DEC);
  *DAYLEN = jmax(0.0001, jmin(0.9999, temp / p_PI));
  double ang1 = temp, ang0 = -temp;
  temp = hafday(L1, DEC);
  double ang7 = temp - L2, ang6 = -temp - L2;
  double ang3 = jmin(ang1, ang7), ang2 = jmax(ang0, ang6);
  double ang8 = 0, ang9 = 0, TWORIS;
  if (ang3 < ang2) { ang2 = 0.; ang3 = 0.; }
  ang6 = ang6 + 2.0 * p_PI;
  if (ang6 < ang1) { ang8 = ang6; ang9 = ang1; TWORIS = 1.; }
  ang7 = ang7 - 2.0 * p_PI;
  if (ang7 > ang0) { ang8 = ang0; ang9 = ang7; TWORIS = 1.0; }
  else TWORIS = 0.0;                       /* overwrites the first branch, SURVEY A.5 #9 */
  double I0SDAY;
  if (TWORIS == 1.0)
    I0SDAY = p_WTOMJ * SCD * (func3(DEC, L2, L1, ang3, ang2) + func3(DEC, L2, L1, ang9, ang8)) / cos(SLOPE);
  else
    I0SDAY = p_WTOMJ * SCD * func3(DEC, L2, L1, ang3, ang2) / cos(SLOPE);
  *I0HDAY = p_WTOMJ * SCD * func3(DEC, 0.0, LAT, ang1, ang0);
  if (*I0HDAY <= 0.0) *SLFDAY = 0.0;
  else *SLFDAY = I0SDAY / *I0HDAY;
}
/* AVAILEN :324-365 */
static void availen(double SLRAD, double ALBEDO, double C1, double C2, double C3, double TA, double EA,
                    double RATIO, double SHEAT, double CR, double LAI, double SAI, double* AA, double* ASUBS) {
  double SOLNET = (1 - ALBEDO) * SLRAD;
  double EFFEM = 1.24 * pow(EA * 10 / (TA + 273.15), 1.0 / 7);
  double NOVERN = (RATIO - C1) / C2;
  if (NOVERN > 1) NOVERN = 1;
  if (NOVERN < 0) NOVERN = 0;
  double CLDCOR = C3 + (1 - C3) * NOVERN;
  double tk = TA + 273.15;
  double LNGNET = (EFFEM - 1) * CLDCOR * p_SIGMA * ((tk * tk) * (tk * tk)); /* x^4 by squaring */
  double RN = SOLNET + LNGNET;
  *AA = RN - SHEAT;
  *ASUBS = RN * exp(-CR * (LAI + SAI)) - SHEAT;
}

/* =========================================================================================
 * PET — src/module_PET.jl
 * ========================================================================================= */
static void esat(double TA, double* ES, double* DELTA) { /* :488-509 */
  *ES = 0.61078 * exp(17.26939 * TA / (TA + 237.3));
  *DELTA = 4098 * *ES / ((TA + 237.3) * (TA + 237.3));
  if (TA < 0) {
    *ES = 0.61078 * exp(21.87456 * TA / (TA + 265.5));
    *DELTA = 5808 * *ES / ((TA + 265.5) * (TA + 265.5));
  }
}
static double pm(double AA, double VPD, double DELTA, double RA, double RC) { /* :1102-1105 */
  return (RA * DELTA * AA + p_CPRHO * VPD) / ((DELTA + p_GAMMA) * RA + p_GAMMA * RC);
}
static void swpe(double AA, double ASUBS, double VPD, double RAA, double RAC, double RAS, double RSC,
                 double RSS, double DELTA, double* PRATE, double* ERATE) { /* :617-657 */
  double RS = (DELTA + p_GAMMA) * RAS + p_GAMMA * RSS;
  double RC = (DELTA + p_GAMMA) * RAC + p_GAMMA * RSC;
  double RA = (DELTA + p_GAMMA) * RAA;
  double CCS = 1 / (1 + RS * RA / (RC * (RS + RA)));
  double CCC = 1 / (1 + RC * RA / (RS * (RC + RA)));
  double PMS = pm(AA, VPD - DELTA * RAS * (AA - ASUBS) / p_CPRHO, DELTA, RAA + RAS, RSS);
  double PMC = pm(AA, VPD - DELTA * RAC * ASUBS / p_CPRHO, DELTA, RAA + RAC, RSC);
  double LE = (CCC * PMC + CCS * PMS);
  double D0 = VPD + RAA * (DELTA * AA - (DELTA + p_GAMMA) * LE) / p_CPRHO;
  *PRATE = p_ETOM * p_WTOMJ * pm(AA - ASUBS, D0, DELTA, RAC, RSC);
  *ERATE = p_ETOM * p_WTOMJ * pm(ASUBS, D0, DELTA, RAS, RSS);
}
static double swge(double AA, double ASUBS, double VPD, double RAA, double RAS, double RSS, double DELTA,
                   double ARATE) { /* :676-694 */
  double LEC = ARATE / (p_ETOM * p_WTOMJ);
  double RS = (DELTA + p_GAMMA) * RAS + p_GAMMA * RSS;
  double RA = (DELTA + p_GAMMA) * RAA;
  double LE = (RS / (RS + RA)) * LEC + (p_CPRHO * VPD + DELTA * RAS * ASUBS + DELTA * RAA * AA) / (RS + RA);
  return p_ETOM * p_WTOMJ * (LE - LEC);
}
static void swgra(double UA, double ZA, double HEIGHT, double Z0, double DISP, double Z0C, double DISPC,
                  double Z0GS, double LWIDTH, double RHOTP, double NN, double LAI, double SAI, double* RAA,
                  double* RAC, double* RAS) { /* :807-840 */
  double USTAR = p_K * UA / (log((ZA - DISP) / Z0));
  double KH = p_K * USTAR * (HEIGHT - DISP);
  *RAS = (HEIGHT * exp(NN) / (NN * KH)) * (exp(-NN * Z0GS / HEIGHT) - exp(-NN * (Z0C + DISPC) / HEIGHT));
  if (*RAS < 1) *RAS = 1.;
  *RAA = log((ZA - DISP) / (HEIGHT - DISP)) / (p_K * USTAR) +
         (HEIGHT / (NN * KH)) * (-1 + exp(NN * (HEIGHT - DISPC - Z0C) / HEIGHT));
  double UH = (USTAR / p_K) * log((HEIGHT - DISP) / Z0);
  double RB = (100 * NN) * pow(LWIDTH / UH, 0.5) / (1 - exp(-NN / 2));
  *RAC = RB / (RHOTP * LAI + p_PI * SAI);
}
static double srsc(double RAD, double TA, double VPD, double LAI, double SAI, double GLMIN, double GLMAX,
                   double R5, double CVPD, double RM, double CR, double TL, double T1, double T2, double TH) {
  /* :1020-1066 */
  double FS = (2 * LAI + SAI) / (2 * LAI);
  double R0, FRINT;
  if (RAD <= 1E-10) { R0 = 0; FRINT = 0; }
  else {
    R0 = RM * R5 / (RM - 2 * R5);
    FRINT = ((RM + R0) / (RM * CR * FS)) * log((R0 + CR * RAD) / (R0 + CR * RAD * exp(-CR * FS * LAI)));
  }
  double FD = 1 / (1 + VPD / CVPD);
  double FT;
  if (TA <= TL) FT = 0;
  else if (TA > TL && TA < T1) { double q = (T1 - TA) / (T1 - TL); FT = 1 - q * q; }
  else if (TA >= T1 && TA <= T2) FT = 1;
  else if (TA > T2 && TA < TH) { double q = (TA - T2) / (TH - T2); FT = 1 - q * q; }
  else FT = 0;
  double GSC = FD * FT * FRINT * (GLMAX - GLMIN) + LAI * GLMIN;
  return 1 / GSC;
}
static double wndadj(double ZA, double DISP, double Z0, double FETCH, double ZW, double Z0W) { /* :543-549 */
  double HIBL = 0.334 * pow(FETCH, 0.875) * pow(Z0W, 0.125);
  return log(HIBL / Z0W) * log((ZA - DISP) / Z0) / (log(HIBL / Z0) * log(ZW / Z0W));
}

/* =========================================================================================
 * EVP — src/module_EVP.jl
 * ========================================================================================= */
/* INTER24 :632-703 */
static void inter24(double RFAL, double PINT, double LAI, double SAI, double FRINTL, double FRINTS,
                    double CINTRL, double CINTRS, int IHD, double INTR, double* RINT, double* IRVP) {
  double CATCH[24], RINTHR[24], IRVPHR[24], INTRNU[24];
  double DTH = 1;
  for (int i = 0; i < 24; i++) {
    if (i < (12 - IHD) || i >= (12 + IHD)) CATCH[i] = 0;
    else CATCH[i] = jmin(1, (FRINTL * LAI + FRINTS * SAI)) * RFAL / (2 * IHD);
  }
  double INTRMX = CINTRL * LAI + CINTRS * SAI;
  INTRNU[0] = INTR;
  for (int i = 0; i < 24; i++) {
    double NEWINT = INTRNU[i] + (CATCH[i] - PINT / 24) * DTH;
    if (NEWINT > 0.0001) {
      IRVPHR[i] = PINT / 24;
      if (NEWINT > INTRMX) RINTHR[i] = IRVPHR[i] + (INTRMX - INTRNU[i]) / DTH;
      else RINTHR[i] = CATCH[i];
    } else {
      IRVPHR[i] = INTRNU[i] / DTH + CATCH[i];
      RINTHR[i] = CATCH[i];
    }
    if (i < 23) INTRNU[i + 1] = INTRNU[i] + (RINTHR[i] - IRVPHR[i]) * DTH;
  }
  double s1 = 0, s2 = 0;
  for (int i = 0; i < 24; i++) s1 += RINTHR[i] * DTH;
  for (int i = 0; i < 24; i++) s2 += IRVPHR[i] * DTH;
  *RINT = s1;
  *IRVP = s2;
}

/* TBYLAYER :301-419 */
static double tbylayer(int J, double PTR, double DISPC, const double* ALPHA, const double* KK,
                       const double* RROOTI, double RXYLEM, const double* PSITI, int N, double PSICR,
                       int NOOUTF, double* ATRANI) {
  int FLAG[MAXN];
  double RI[MAXN];
  for (int i = 0; i < N; i++) {
    if (RROOTI[i] > 1E+15) FLAG[i] = 1;
    else if (NOOUTF && PSITI[i] / 1000. <= PSICR) FLAG[i] = 1;
    else FLAG[i] = 0;
    RI[i] = 0; ATRANI[i] = 0;
  }
  double ATR = 0.0;
  for (;;) {
    int NEGFLAG = 0;
    double SUM = 0;
    for (int i = 0; i < N; i++) {
      if (FLAG[i] == 0) { RI[i] = RROOTI[i] + ALPHA[i] / KK[i]; SUM = SUM + 1.0 / RI[i]; }
      else ATRANI[i] = 0.0;
    }
    double RT;
    if (SUM < 1E-20) { ATR = 0.; break; }
    else RT = 1.0 / SUM;
    double PSIT = 0;
    for (int i = 0; i < N; i++)
      if (FLAG[i] == 0) PSIT = PSIT + RT * PSITI[i] / RI[i];
    double SUPPLY = (PSIT / 1000 - PSICR - p_RHOWG * DISPC) / (RT + RXYLEM);
    if (J == 1) {
      double R = (2 / p_PI) * (SUPPLY / PTR);
      if (R <= 0) ATR = 0.;
      else if (R < 1) ATR = PTR * (1 + R * acos(R) - sin(acos(R)));
      else ATR = PTR;
    } else {
      if ((SUPPLY <= 0) || (PTR <= 0)) ATR = 0.;
      else ATR = jmin(SUPPLY, PTR);
    }
    for (int i = 0; i < N; i++) {
      if (FLAG[i] == 1) ATRANI[i] = 0;
      else {
        ATRANI[i] = ((PSITI[i] - PSIT) / 1000 + RT * ATR) / RI[i];
        if (ATRANI[i] < -0.000001) NEGFLAG = 1;
      }
    }
    if (NOOUTF && NEGFLAG) {
      int IDEL = -1;
      double TRMIN = 0;
      for (int i = 0; i < N; i++)
        if (ATRANI[i] < TRMIN) { TRMIN = ATRANI[i]; IDEL = i; }
      FLAG[IDEL] = 1;
    } else break;
  }
  return ATR;
}

/* =========================================================================================
 * SNO — src/module_SNO.jl
 * ========================================================================================= */
static double snofrac(double TMAX, double TMIN, double RSTEMP) { /* :98-107 */
  if (TMIN >= RSTEMP) return 0.0;
  else if (TMAX < RSTEMP) return 1.0;
  return 1.0 - (TMAX - RSTEMP) / (TMAX - TMIN);
}
/* SNOWPACK :311-493 */
static void snowpack(double RTHR, double STHR, double PSNVP, double SNOEN, double* CC, double* SNOW,
                     double* SNOWLQ, double DTP, double TA, double MAXLQF, double GRDMLT, double* RSNO,
                     double* SNVP, double* SMLT) {
  double u_CC = *CC, u_SNOW = *SNOW, u_SNOWLQ = *SNOWLQ;
  double aux_du_RSNO = 0, aux_du_SNVP, aux_du_SMLT;
  u_SNOW = u_SNOW + STHR * DTP;
  u_CC = u_CC + (p_CVICE * jmax(-TA, 0.) * STHR) * DTP;
  if (u_CC > 0. && u_SNOWLQ > 0.) {
    if (u_CC > u_SNOWLQ * p_LF) { u_CC = u_CC - u_SNOWLQ * p_LF; u_SNOWLQ = 0.; }
    else { u_SNOWLQ = u_SNOWLQ - u_CC / p_LF; u_CC = 0.; }
  }
  double FRAC = (GRDMLT + PSNVP) * DTP / u_SNOW;
  if (FRAC < 1) {
    aux_du_SMLT = GRDMLT;
    aux_du_SNVP = PSNVP;
    u_SNOW = u_SNOW * (1. - FRAC);
    u_SNOWLQ = u_SNOWLQ * (1. - FRAC);
    u_CC = u_CC * (1. - FRAC);
  } else {
    aux_du_SMLT = GRDMLT / FRAC;
    aux_du_SNVP = PSNVP / FRAC;
    u_SNOW = 0.; u_SNOWLQ = 0.; u_CC = 0.;
  }
  if (u_SNOW == 0.) aux_du_RSNO = 0.;
  if (u_SNOW > 0) {
    double EQEN = DTP * (SNOEN + RTHR * jmax(TA, 0) * p_CVLQ) / p_LF;
    if (EQEN <= 0) {
      double NMLT = -EQEN;
      if (NMLT < u_SNOWLQ) { u_CC = 0.; u_SNOWLQ = u_SNOWLQ - NMLT; }
      else {
        NMLT = NMLT - u_SNOWLQ;
        u_SNOWLQ = 0.;
        u_CC = u_CC + NMLT * p_LF;
        u_CC = jmin(u_CC, -TA * u_SNOW * p_CVICE);
      }
    } else {
      if (EQEN * p_LF < u_CC || TA < 0) {
        if (TA < 0) { u_CC = jmax(u_CC - EQEN * p_LF, -TA * u_SNOW * p_CVICE); u_SNOWLQ = 0.; }
        else { u_CC = u_CC - EQEN * p_LF; u_SNOWLQ = 0.; }
      } else {
        EQEN = EQEN - u_CC / p_LF;
        if (EQEN <= MAXLQF * u_SNOW - u_SNOWLQ) { u_SNOWLQ = u_SNOWLQ + EQEN; u_CC = 0.; }
        else {
          EQEN = EQEN - (MAXLQF * u_SNOW - u_SNOWLQ);
          if (u_SNOW * (1. - MAXLQF) > EQEN) {
            aux_du_SMLT = aux_du_SMLT + (EQEN / DTP) / (1. - MAXLQF);
            u_SNOW = u_SNOW - EQEN / (1. - MAXLQF);
            u_SNOWLQ = MAXLQF * u_SNOW;
            u_CC = 0.;
          } else {
            aux_du_SMLT = aux_du_SMLT + u_SNOW / DTP;
            u_SNOW = 0.; u_SNOWLQ = 0.; u_CC = 0.;
          }
        }
      }
    }
    if (RTHR == 0 || u_SNOW == 0) aux_du_RSNO = 0.;
    else {
      double RIN = RTHR * DTP;
      if (u_CC > 0) {
        if (u_CC > RIN * p_LF) { u_CC = u_CC - RIN * p_LF; aux_du_RSNO = RTHR; u_SNOW = u_SNOW + RIN; }
        else {
          u_SNOW = u_SNOW + u_CC / p_LF;
          aux_du_RSNO = (u_CC / p_LF) / DTP;
          u_CC = 0.;
          RIN = RIN - aux_du_RSNO * DTP;
          if (RIN < MAXLQF * u_SNOW / (1. - MAXLQF)) {
            u_SNOWLQ = RIN;
            aux_du_RSNO = aux_du_RSNO + RIN / DTP;
            u_SNOW = u_SNOW + RIN;
          } else {
            u_SNOWLQ = MAXLQF * u_SNOW / (1. - MAXLQF);
            aux_du_RSNO = aux_du_RSNO + u_SNOWLQ / DTP;
            u_SNOW = u_SNOW + u_SNOWLQ;
          }
        }
      } else {
        if (u_SNOWLQ >= MAXLQF * u_SNOW) aux_du_RSNO = 0.;
        else {
          double ALQ = MAXLQF * u_SNOW - u_SNOWLQ;
          if (RIN < ALQ) { aux_du_RSNO = RTHR; u_SNOWLQ = u_SNOWLQ + RIN; u_SNOW = u_SNOW + RIN; }
          else {
            aux_du_RSNO = (ALQ / (1. - MAXLQF)) / DTP;
            u_SNOW = u_SNOW + aux_du_RSNO * DTP;
            u_SNOWLQ = MAXLQF * u_SNOW;
          }
        }
      }
    }
  }
  *CC = u_CC; *SNOW = u_SNOW; *SNOWLQ = u_SNOWLQ;
  *RSNO = aux_du_RSNO; *SNVP = aux_du_SNVP; *SMLT = aux_du_SMLT;
}

/* =========================================================================================
 * State-vector offsets (src/func_DiffEq_definition_u0.jl:42-73)
 * ========================================================================================= */
typedef struct {
  int k, SWATI, SWATI18, SWATI2, GWAT, INTS, INTR, SNOW, RWU, XYLEM, CC, SNOWLQ, TRANI, AUX, ACC, L;
} sidx_t;
static sidx_t make_idx(int N, int iso) {
  sidx_t s;
  s.k = iso ? 3 : 1;
  int o = 0;
  s.SWATI = 0; s.SWATI18 = N; s.SWATI2 = 2 * N;
  o += s.k * N;
  s.GWAT = o; o += s.k;
  s.INTS = o; o += s.k;
  s.INTR = o; o += s.k;
  s.SNOW = o; o += s.k;
  s.RWU = o; o += s.k;
  s.XYLEM = o; o += s.k;
  s.CC = o; s.SNOWLQ = o + 1; o += 2;
  s.TRANI = o; o += s.k * N;
  s.AUX = o; o += 3 * N;
  s.ACC = o; o += LWFB_NACCUM;
  s.L = o;
  return s;
}

/* relative root density of the day: HammelKennel_lateral_rootgrowth,
 * src/func_DiffEq_definition_p.jl:606-628, evaluated at integer days (Constant{Next} on the day
 * grid, src/func_discretize_soil_domain.jl:632-634) with age = AGE0 + t/365 (:856). */
static void relden_of_day(const model_t* m, double t, double* RELDEN) {
  int N = m->N;
  if (SC(m, RELDEN_CONST) != 0.0) {
    for (int i = 0; i < N; i++) RELDEN[i] = LY(m, TINI, i) > 1e19 ? 0.0 : LY(m, RELDEN, i);
    return;
  }
  double t_y = SC(m, AGE0) + t / 365, per = SC(m, RGROPER), rl0 = SC(m, RL0);
  for (int i = 0; i < N; i++) {
    double ts = LY(m, TINI, i), fin = LY(m, RELDEN, i);
    if (t_y < ts) RELDEN[i] = 0.0;
    else if (t_y <= ts + per) RELDEN[i] = rl0 * pow(fin / rl0, (t_y - ts) / per);
    else RELDEN[i] = fin;
  }
}

/* =========================================================================================
 * Daily callback: LWFBrook90R_updateAmounts_INTS_INTR_SNOW_CC_SNOWLQ!
 * src/func_DiffEq_definition_cb.jl:60-302 with MSBSETVARS (src/func_MSB_functions.jl:31-132),
 * MSBDAYNIGHT (:211-324), MSBDAYNIGHT_postprocess (:338-364), MSBPREINT (:382-507).
 * ========================================================================================= */
static void cb_amounts(model_t* m, double* u, const sidx_t* ix, int day) {
  int N = m->N;
  double W[MAXN], psim[MAXN], psiti[MAXN], theta[MAXN], kk[MAXN];
  double u_INTS = u[ix->INTS], u_INTR = u[ix->INTR], u_SNOW = u[ix->SNOW], u_CC = u[ix->CC],
         u_SNOWLQ = u[ix->SNOWLQ];
  double* swati = u + ix->SWATI;
  swchek(m, swati);                                  /* cb.jl:131 */
  soilvar(m, swati, W, psim, psiti, theta, kk);      /* cb.jl:134-135 (locals) */

  double TMAX = FC(m, day, TMAX), TMIN = FC(m, day, TMIN), EA = FC(m, day, VAPPRES),
         UW = FC(m, day, WIND), SOLRAD = FC(m, day, GLOBRAD), PREC = FC(m, day, PREC),
         DOY = FC(m, day, DOY);
  /* canopy of the day: make_absolute_from_relative, src/func_DiffEq_definition_p.jl:816-833 */
  double DENSEF = FC(m, day, DENSEF_REL) * SC(m, DENSEF_BASELINE);
  double HEIGHT = FC(m, day, HEIGHT_REL) * SC(m, HEIGHT_BASELINE);
  double LAI = FC(m, day, LAI_REL) * SC(m, MAXLAI);
  double SAI = FC(m, day, SAI_REL) * SC(m, SAI_BASELINE);
  double RELDEN[MAXN];
  relden_of_day(m, (double)day, RELDEN);

  /* ---- MSBSETVARS ---- */
  double DAYLEN, I0HDAY, SLFDAY;
  sunds(SC(m, LAT), SC(m, ESLOPE), DOY, SC(m, L1), SC(m, L2), &DAYLEN, &I0HDAY, &SLFDAY);
  /* CANOPY_timeEvolution  module_PET.jl:193-211 */
  double SAIeff = DENSEF * SAI;
  double RTLENeff = DENSEF * SC(m, MXRTLN);
  double KPL = jmax(DENSEF * SC(m, MXKPL), 1E-8);
  double RPLANT = 1 / KPL;
  double LAIeff0 = DENSEF * LAI;
  double SNOFRC = snofrac(TMAX, TMIN, SC(m, RSTEMP));
  /* PLNTRES  module_EVP.jl:144-190 */
  double D[MAXN], RTFRAC[MAXN], RROOTI[MAXN], ALPHA[MAXN];
  double sumRD = 0;
  for (int i = 0; i < N; i++) D[i] = LY(m, THICK, i) * (1 - LY(m, STONEF, i));
  for (int i = 0; i < N; i++) sumRD += RELDEN[i] * D[i];
  for (int i = 0; i < N; i++) RTFRAC[i] = RELDEN[i] * D[i] / sumRD;
  double RXYLEM = SC(m, FXYLEM) * RPLANT;
  for (int i = 0; i < N; i++) {
    if (RELDEN[i] < 0.00001 || RTLENeff < 0.1) { RROOTI[i] = 1E+20; ALPHA[i] = 1E+20; }
    else {
      RROOTI[i] = (RPLANT - RXYLEM) / RTFRAC[i];
      double RTDENI = RTFRAC[i] * 0.001 * RTLENeff / D[i];
      double DELT = p_PI * (SC(m, RTRAD) * SC(m, RTRAD)) * RTDENI;
      ALPHA[i] = (1 / (8 * p_PI * RTDENI)) * (DELT - 3 - 2 * (log(DELT)) / (1 - DELT));
      ALPHA[i] = ALPHA[i] * 0.001 * p_RHOWG / D[i];
    }
  }
  /* CANOPY_snowCover  module_PET.jl:230-258 */
  double SNODEP = 0.001 * u_SNOW / SC(m, SNODEN);
  double HSNO = jmax(HEIGHT - SNODEP, 0);
  double RATIOc = HSNO / HEIGHT;
  double HEIGHTeff = jmax(HSNO, 0.01);
  double LAIeff = jmax(RATIOc * LAIeff0, 0.00001);
  /* ROUGH  module_PET.jl:325-366 */
  double Z0GS = (u_SNOW > 0) ? SC(m, Z0S) : SC(m, Z0G);
  double Z0C, DISPC, Z0, DISP, ZA;
  {
    double HR = SC(m, HR), HS = SC(m, HS), CZR = SC(m, CZR), CZS = SC(m, CZS);
    if (HEIGHTeff >= HR) Z0C = CZR * HEIGHTeff;
    else if (HEIGHTeff <= HS) Z0C = CZS * HEIGHTeff;
    else Z0C = CZS * HS + (CZR * HR - CZS * HS) * (HEIGHTeff - HS) / (HR - HS);
    DISPC = HEIGHTeff - Z0C / 0.3;
    if (Z0GS > Z0C) Z0GS = Z0C;
    double RATIO = (LAIeff + SAIeff) / (SC(m, LPC) + 0.035 * HEIGHTeff);
    if (RATIO >= 1) { Z0 = Z0C; DISP = DISPC; }
    else {
      double e = -1 + exp(0.909 - 3.03 * Z0C / HEIGHTeff);
      double XX = RATIO * ((e * e) * (e * e));
      DISP = 1.1 * HEIGHTeff * log(1.0 + pow(XX, 0.25));
      Z0 = jmin(0.3 * (HEIGHTeff - DISP), Z0GS + 0.3 * HEIGHTeff * pow(XX, 0.5));
    }
    ZA = HEIGHTeff + SC(m, ZMINH);
  }
  double SHEAT = 0.;
  /* WEATHER  module_PET.jl:429-480 */
  double SOLRADC, TA, TADTM, TANTM, UA, UADTM, UANTM;
  {
    if (SOLRAD < 0.001) SOLRADC = 0.550 * I0HDAY;
    else if (SOLRAD > I0HDAY) SOLRADC = 0.99 * I0HDAY;
    else SOLRADC = SOLRAD;
    TA = (TMAX + TMIN) / 2;
    TADTM = TA + ((TMAX - TMIN) / (2 * p_PI * DAYLEN)) * sin(p_PI * DAYLEN);
    TANTM = TA - ((TMAX - TMIN) / (2 * p_PI * (1 - DAYLEN))) * sin(p_PI * DAYLEN);
    double uw = UW;
    if (uw == 0) uw = 3.0;
    uw = jmax(uw, 0.2);
    UA = uw * wndadj(ZA, DISP, Z0, SC(m, FETCH), SC(m, ZW), SC(m, Z0W));
    UADTM = UA / (DAYLEN + (1 - DAYLEN) * SC(m, WNDRAT));
    UANTM = SC(m, WNDRAT) * UADTM;
  }
  double TSNOW, PSNVP, ALBEDO;
  if (u_SNOW > 0) {
    TSNOW = -u_CC / (p_CVICE * u_SNOW);
    /* SNOVAP  module_SNO.jl:166-183 */
    double ESNOW, dummy;
    if (TSNOW > -0.1) ESNOW = 0.61;
    else esat(jmin(TA, TSNOW), &ESNOW, &dummy);
    double RAA, RAC, RAS;
    swgra(UA, ZA, HEIGHTeff, Z0, DISP, Z0C, DISPC, Z0GS, SC(m, LWIDTH), SC(m, RHOTP), SC(m, NN), LAIeff, SAIeff,
          &RAA, &RAC, &RAS);
    PSNVP = (p_WTOMJ / p_LS) * (p_CPRHO / p_GAMMA) * (ESNOW - EA) / (RAA + RAS);
    PSNVP = SC(m, KSNVP) * PSNVP;
    ALBEDO = SC(m, ALBSN);
  } else { TSNOW = 0.; PSNVP = 0.; ALBEDO = SC(m, ALB); }
  /* SNOENRGY  module_SNO.jl:223-234 */
  double SNOEN;
  if (TA <= 0) SNOEN = SC(m, CCFAC) * 2.0 * DAYLEN * (TA - TSNOW);
  else SNOEN = SC(m, MELFAC) * 2.0 * DAYLEN * TA * exp(-SC(m, SAIMLT) * SAIeff) * exp(-SC(m, LAIMLT) * LAIeff) * SLFDAY;
  double RSS;
  if (u_SNOW > 0) RSS = 0.;
  else { /* FRSS  module_PET.jl:873-888 */
    if (SC(m, RSSA) < 0.0001) RSS = 1e20;
    else RSS = SC(m, RSSA) * pow(jmax(psim[0] / LY(m, PSIF, 0), 1.0), SC(m, RSSB));
    if (RSS < 0.000001) m->status = LWFB_ST_RSS; /* reference raises, MSB:114-119 */
  }

  /* ---- MSBDAYNIGHT ---- */
  double PTR[2], GER[2], PIR[2], GIR[2], PGER[2], ATRI[2][MAXN];
  for (int J = 1; J <= 2; J++) {
    double SLRAD, TAJ, UAJ;
    if (J == 1) { SLRAD = SLFDAY * SOLRADC / (p_WTOMJ * DAYLEN); TAJ = TADTM; UAJ = UADTM; }
    else { SLRAD = 0; TAJ = TANTM; UAJ = UANTM; }
    double cloud = SOLRADC / I0HDAY;
    double AA, ASUBS, ES, DELTA, RAA, RAC, RAS, RSC, dmy;
    availen(SLRAD, ALBEDO, SC(m, C1), SC(m, C2), SC(m, C3), TAJ, EA, cloud, SHEAT, SC(m, CR), LAIeff, SAIeff, &AA, &ASUBS);
    esat(TAJ, &ES, &DELTA);
    double VPD = ES - EA;
    swgra(UAJ, ZA, HEIGHTeff, Z0, DISP, Z0C, DISPC, Z0GS, SC(m, LWIDTH), SC(m, RHOTP), SC(m, NN), LAIeff, SAIeff,
          &RAA, &RAC, &RAS);
    if (J == 1)
      RSC = srsc(SLRAD, TA, VPD, LAIeff, SAIeff, SC(m, GLMIN), SC(m, GLMAX), SC(m, R5), SC(m, CVPD), SC(m, RM),
                 SC(m, CR), SC(m, TL), SC(m, T1), SC(m, T2), SC(m, TH));
    else RSC = 1 / (SC(m, GLMIN) * LAIeff);
    swpe(AA, ASUBS, VPD, RAA, RAC, RAS, RSC, RSS, DELTA, &PTR[J - 1], &GER[J - 1]);
    swpe(AA, ASUBS, VPD, RAA, RAC, RAS, 0, RSS, DELTA, &PIR[J - 1], &GIR[J - 1]);
    swpe(AA, ASUBS, VPD, RAA, RAC, RAS, RSC, 0, DELTA, &dmy, &PGER[J - 1]);
    if (PTR[J - 1] > 0.001) {
      double ATR = tbylayer(J, PTR[J - 1], DISPC, ALPHA, kk, RROOTI, RXYLEM, psiti, N, SC(m, PSICR),
                            SC(m, NOOUTF) != 0.0, ATRI[J - 1]);
      if (ATR < PTR[J - 1]) GER[J - 1] = swge(AA, ASUBS, VPD, RAA, RAS, RSS, DELTA, ATR);
    } else {
      PTR[J - 1] = 0;
      for (int i = 0; i < N; i++) ATRI[J - 1][i] = 0;
      GER[J - 1] = swge(AA, ASUBS, VPD, RAA, RAS, RSS, DELTA, 0);
    }
  }
  /* ---- MSBDAYNIGHT_postprocess ---- */
  double PTRAN = (PTR[0] * DAYLEN + PTR[1] * (1 - DAYLEN));
  double GEVP = (GER[0] * DAYLEN + GER[1] * (1 - DAYLEN));
  double PINT = (PIR[0] * DAYLEN + PIR[1] * (1 - DAYLEN));
  double GIVP = (GIR[0] * DAYLEN + GIR[1] * (1 - DAYLEN));
  for (int i = 0; i < N; i++) m->TRANI[i] = (ATRI[0][i] * DAYLEN + ATRI[1][i] * (1 - DAYLEN));

  /* ---- MSBPREINT ---- */
  double DTP = 1.0;
  double SFAL = SNOFRC * PREC, RFAL = PREC - SFAL;
  int IHD = (int)FC(m, day, IHD);
  double SINT, ISVP, RINT, IRVP;
  if (PINT < 0 && TA > 0)
    inter24(SFAL, 0, LAIeff, SAIeff, SC(m, FSINTL), SC(m, FSINTS), SC(m, CINTSL), SC(m, CINTSS), IHD, u_INTS, &SINT, &ISVP);
  else
    inter24(SFAL, PINT, LAIeff, SAIeff, SC(m, FSINTL), SC(m, FSINTS), SC(m, CINTSL), SC(m, CINTSS), IHD, u_INTS, &SINT, &ISVP);
  inter24(RFAL, PINT - ISVP, LAIeff, SAIeff, SC(m, FRINTL), SC(m, FRINTS), SC(m, CINTRL), SC(m, CINTRS), IHD, u_INTR, &RINT, &IRVP);
  double WETFR = jmin(1.0, (IRVP + ISVP) / PINT);
  double RTHR = RFAL - RINT, STHR = SFAL - SINT;
  PTRAN = (1.0 - WETFR) * PTRAN;
  for (int i = 0; i < N; i++) m->TRANI[i] = (1.0 - WETFR) * m->TRANI[i];
  double SLVP, RSNO, SNVP, SMLT;
  if (u_SNOW <= 0 && STHR <= 0) SLVP = WETFR * GIVP + (1.0 - WETFR) * GEVP;
  else SLVP = 0.0;
  double nSNOW = u_SNOW, nCC = u_CC, nSNOWLQ = u_SNOWLQ;
  if (u_SNOW <= 0 && STHR <= 0) { RSNO = 0.0; SNVP = 0.0; SMLT = 0.0; }
  else {
    if (u_SNOW <= 0 && STHR > 0) { nCC = 0.0; nSNOWLQ = 0.0; }
    snowpack(RTHR, STHR, PSNVP, SNOEN, &nCC, &nSNOW, &nSNOWLQ, DTP, TA, SC(m, MAXLQF), SC(m, GRDMLT), &RSNO, &SNVP, &SMLT);
  }
  double RNET = RTHR - RSNO;

  /* ---- state and cache writes cb.jl:227-289 ---- */
  m->RNET = RNET; m->SMLT = SMLT; m->SLVP = SLVP; m->TA = TA;
  u[ix->INTS] = u_INTS + (SINT - ISVP) * DTP;
  u[ix->INTR] = u_INTR + (RINT - IRVP) * DTP;
  u[ix->SNOW] = nSNOW;
  u[ix->CC] = nCC;
  u[ix->SNOWLQ] = nSNOWLQ;
  double sumT = 0;
  for (int i = 0; i < N; i++) sumT += m->TRANI[i];
  for (int i = 0; i < N; i++) u[ix->TRANI + i] = DTP * m->TRANI[i];
  u[ix->RWU] = DTP * sumT;
  double* a = u + ix->ACC;
  a[LWFB_A_CUM_D_PREC] = DTP * (RFAL + SFAL);
  a[LWFB_A_CUM_D_RFAL] = DTP * RFAL;
  a[LWFB_A_CUM_D_SFAL] = DTP * SFAL;
  a[LWFB_A_CUM_D_RINT] = DTP * RINT;
  a[LWFB_A_CUM_D_SINT] = DTP * SINT;
  a[LWFB_A_CUM_D_RSNO] = DTP * RSNO;
  a[LWFB_A_CUM_D_RNET] = DTP * (RFAL - RINT - RSNO);
  a[LWFB_A_CUM_D_SMLT] = DTP * SMLT;
  a[LWFB_A_CUM_D_EVAP] = DTP * (IRVP + ISVP + SNVP + SLVP + sumT);
  a[LWFB_A_CUM_D_TRAN] = DTP * sumT;
  a[LWFB_A_CUM_D_IRVP] = DTP * IRVP;
  a[LWFB_A_CUM_D_ISVP] = DTP * ISVP;
  a[LWFB_A_CUM_D_SLVP] = DTP * SLVP;
  a[LWFB_A_CUM_D_SNVP] = DTP * SNVP;
  a[LWFB_A_CUM_D_PINT] = DTP * PINT;
  a[LWFB_A_CUM_D_PTRAN] = DTP * PTRAN;
  a[LWFB_A_CUM_D_PSLVP] = DTP * (PGER[0] * DAYLEN + PGER[1] * (1 - DAYLEN));
  a[LWFB_A_FLOW] = 0; a[LWFB_A_SEEP] = 0; a[LWFB_A_SRFL] = 0; a[LWFB_A_SLFL] = 0;
  a[LWFB_A_BYFL] = 0; a[LWFB_A_DSFL] = 0; a[LWFB_A_GWFL] = 0; a[LWFB_A_VRFLN] = 0;
  a[LWFB_A_CUM_D_RTHR] = DTP * (RFAL - RINT);
  a[LWFB_A_CUM_D_STHR] = DTP * (SFAL - SINT);
}

/* δ of precipitation at time t within [day, day+1]:  Gridded(Constant{Next}) on sample day + 0.5
 * (src/func_DiffEq_definition_p.jl:735-736): the value of row r holds for t in (r-0.5, r+0.5]. */
static void prec_delta(const model_t* m, double t, double* d18, double* d2) {
  int r = (int)ceil(t - 0.5);
  if (r < 0) r = 0;
  if (r > m->n_rows - 1) r = m->n_rows - 1;
  *d18 = FC(m, r, D18O);
  *d2 = FC(m, r, D2H);
}

/* LWFBrook90R_updateIsotopes_INTS_INTR_SNOW!  src/func_DiffEq_definition_cb.jl:304-530 (live part
 * :380-387, :520-525): δ of INTS, INTR, SNOW := δ of precipitation at t */
static void cb_deltas(model_t* m, double* u, const sidx_t* ix, double t) {
  if (!m->iso) return;
  double d18, d2;
  prec_delta(m, t, &d18, &d2);
  u[ix->INTS + 1] = d18; u[ix->INTS + 2] = d2;
  u[ix->INTR + 1] = d18; u[ix->INTR + 2] = d2;
  u[ix->SNOW + 1] = d18; u[ix->SNOW + 2] = d2;
}

/* src/module_ISO.jl:36-37, :53-55 */
static const double R18 = 2005.2e-6, R2H = 155.76e-6;
static double d_to_x(double d, double R) { return 1 / (1 + 1 / (R * (DIVC(d, 1000.0) + 1))); }
static double x_to_d(double x, double R) { return (1 / (R * (1 / x - 1)) - 1) * 1000; }
static double dxdt_to_dddt(double dxdt, double x, double R) {
  return DIVC(dxdt * 1, R) * 1 / ((x - 1) * (x - 1)) * 1000;
}

/* LWFBrook90R_updateIsotopes_GWAT_SWAT_AdvecDiff!  src/func_DiffEq_definition_cb.jl:635-1250.
 * Explicit Euler step of δ18O, δ2H of SWATI, GWAT, and RWU / XYLEM mixing, using the flux caches
 * left by the last f call. */
static void cb_ade(model_t* m, double* u, const double* uprev, const sidx_t* ix, double t, double tprev) {
  if (!m->iso) return;
  int N = m->N;
  double W[MAXN], psim[MAXN], psiti[MAXN], th1[MAXN], kk[MAXN];
  const double* S1 = u + ix->SWATI;
  double G1 = u[ix->GWAT];
  soilvar(m, S1, W, psim, psiti, th1, kk);                       /* cb.jl:675 (only θ is used) */
  double dprec18, dprec2;
  prec_delta(m, t, &dprec18, &dprec2);
  double dt = t - tprev;
  double Tk = m->TA + 273.15;
  double Lam = SC(m, DISPERSIVITY);
  double e = exp(-535400 / (Tk * Tk) + 1393.3 / Tk + 2.1876);
  double D0[2] = {0.96691 * 1e-9 * e * 3600 * 24, 0.98331 * 1e-9 * e * 3600 * 24};
  const double R[2] = {R18, R2H};
  double dprec[2] = {dprec18, dprec2};
  double sumT = 0;
  for (int i = 0; i < N; i++) sumT += m->TRANI[i];
  for (int s = 0; s < 2; s++) {
    int o = (s + 1);
    const double* dk = uprev + ix->SWATI + o * N;   /* δ at the start of the step */
    double* dnew = u + ix->SWATI + o * N;
    double C[MAXN], Dd[MAXN], dC[MAXN], dd[MAXN], Ctr[MAXN];
    double Cxy = d_to_x(u[ix->XYLEM + o], R[s]);
    double dSLVP = dnew[0];                          /* cb.jl:814-815 (before the update) */
    for (int i = 0; i < N; i++) {
      double q = DIVC(m->VRFLI[i], 1000.0);
      double ths = LY(m, THSAT, i);
      double tau = DIVC(pow(th1[i], 7.0 / 3), ths * ths);
      Dd[i] = D0[s] * tau * th1[i] + Lam * fabs(q);
      C[i] = d_to_x(dk[i], R[s]);
    }
    double Cinf = d_to_x(dprec[s], R[s]);
    double Cslvp = d_to_x(dSLVP, R[s]);
    for (int i = 0; i < N; i++) Ctr[i] = (m->TRANI[i] < 0) ? Cxy : C[i];
    for (int i = 0; i < N; i++) {
      double diff_upp = (i == 0) ? 0.0 * 0.0 : ((Dd[i - 1] + Dd[i]) / 2) * DIVC(C[i] - C[i - 1], LY(m, DZ, i - 1));
      double diff_low = (i == N - 1) ? 0.0 * 0.0 : ((Dd[i] + Dd[i + 1]) / 2) * DIVC(C[i + 1] - C[i], LY(m, DZ, i));
      double qC_upp = (i == 0) ? 0.0 : m->VRFLI[i - 1] * C[i - 1];
      double qC_low = m->VRFLI[i] * C[i];
      double slv = (i == 0) ? Cslvp : 0.0;
      double inner = -diff_upp * 1000 + diff_low * 1000 + qC_upp;
      inner = inner - qC_low;
      inner = inner + m->INFLI[i] * Cinf;
      inner = inner - m->DSFLI[i] * C[i];
      inner = inner - m->SLVP * slv;
      inner = inner - m->TRANI[i] * Ctr[i];
#if defined(LWFB_ORACLE_ELEM) && !defined(LWFB_ORACLE_NORCP)
      dC[i] = -C[i] * (1 / S1[i]) * m->NTFLI[i] + 1 / S1[i] * inner;
#else
      dC[i] = -C[i] / S1[i] * m->NTFLI[i] + 1 / S1[i] * inner;
#endif
    }
    /* GWAT cb.jl:1152-1188 */
    double ddG;
    double VN = m->VRFLI[N - 1];
    if ((G1 == 0) && (VN == 0)) ddG = NAN;
    else if ((G1 == 0) && (VN > 0)) ddG = 0;
    else {
      double CG = d_to_x(u[ix->GWAT + o], R[s]);
      double dCG = VN / G1 * (d_to_x(dnew[N - 1], R[s]) - CG);
      ddG = dxdt_to_dddt(dCG, CG, R[s]);
    }
    for (int i = 0; i < N; i++) dd[i] = dxdt_to_dddt(dC[i], C[i], R[s]);
    u[ix->GWAT + o] += dt * ddG;
    for (int i = 0; i < N; i++) dnew[i] += dt * dd[i];
    /* RWU / XYLEM cb.jl:1218-1243 */
    double sw = 0;
    for (int i = 0; i < N; i++) sw += Ctr[i] * m->TRANI[i];
    double dRWU = x_to_d(sw / sumT, R[s]);
    u[ix->RWU + o] = dRWU;
    double dxy = u[ix->XYLEM + o];
    double ddxy = isnan(dRWU) ? 0.0 : sumT / SC(m, VXYLEM) * (dRWU - dxy);
    u[ix->XYLEM + o] += dt * ddxy;
  }
}

/* LWFBrook90R_check_balance_errors!  src/func_DiffEq_definition_cb.jl:1255-1314 */
static void cb_balance(model_t* m, double* u, const double* uprev, const sidx_t* ix) {
  int N = m->N;
  double* a = u + ix->ACC;
  double old_SWAT = uprev[ix->ACC + LWFB_A_STORAGESWAT], old_W = uprev[ix->ACC + LWFB_A_STORAGEWATER];
  double S = 0;
  for (int i = 0; i < N; i++) S += u[ix->SWATI + i];
  double Wt = u[ix->INTR] + u[ix->INTS] + u[ix->SNOW] + S + u[ix->GWAT];
  double qin = 0 + (a[LWFB_A_SLFL] - a[LWFB_A_BYFL]);
  double qout = a[LWFB_A_VRFLN] + a[LWFB_A_DSFL] + a[LWFB_A_CUM_D_TRAN] + a[LWFB_A_CUM_D_SLVP];
  double qin_t = a[LWFB_A_CUM_D_PREC];
  double qout_t = a[LWFB_A_CUM_D_EVAP] + a[LWFB_A_FLOW] + a[LWFB_A_SEEP];
  a[LWFB_A_BALERD_SWAT] = (qin - qout) - (S - old_SWAT);
  a[LWFB_A_BALERD_TOTAL] = (qin_t - qout_t) - (Wt - old_W);
  a[LWFB_A_STORAGESWAT] = S;
  a[LWFB_A_STORAGEWATER] = Wt;
}

/* =========================================================================================
 * SOLVER — restated from the published algorithm of OrdinaryDiffEq.jl v6 (Tsit5, PIController,
 * tstop handling) and DiffEqBase.fastpow; call sites src/LWFBrook90.jl:552-567, :643.
 * ========================================================================================= */

/* DiffEqBase.fastpow: Float32 log2 approximation (Goldberg, "fast approximate logarithms",
 * table 2 line 8) followed by a Float32 exp2 (openlibm s_exp2f table algorithm). */
static float fastlog2f(float x) {
  const float a = 0.338953f, b = 2.198599f, c = 1.523692f;
  uint32_t ux1i;
  memcpy(&ux1i, &x, 4);
  int32_t ex = (int32_t)((ux1i & 0x7F800000u) >> 23);
  uint32_t greater = ux1i & 0x00400000u;
  float signif, fexp, lg2;
  if (greater != 0) {
    uint32_t ux2i = (ux1i & 0x007FFFFFu) | 0x3f000000u;
    memcpy(&signif, &ux2i, 4);
    fexp = (float)ex - 126.0f;
    signif = signif - 1.0f;
    lg2 = fexp + signif * (a * signif + b) / (signif + c);
  } else {
    uint32_t ux2i = (ux1i & 0x007FFFFFu) | 0x3f800000u;
    memcpy(&signif, &ux2i, 4);
    fexp = (float)ex - 127.0f;
    signif = signif - 1.0f;
    lg2 = fexp + signif * (a * signif + b) / (signif + c);
  }
  return lg2;
}
static float exp2f_tbl(float x) {
  static const double tbl[16] = {
      0x1.6a09e667f3bcdp-1, 0x1.7a11473eb0187p-1, 0x1.8ace5422aa0dbp-1, 0x1.9c49182a3f090p-1,
      0x1.ae89f995ad3adp-1, 0x1.c199bdd85529cp-1, 0x1.d5818dcfba487p-1, 0x1.ea4afa2a490dap-1,
      0x1.0000000000000p+0, 0x1.0b5586cf9890fp+0, 0x1.172b83c7d517bp+0, 0x1.2387a6e756238p+0,
      0x1.306fe0a31b715p+0, 0x1.3dea64c123422p+0, 0x1.4bfdad5362a27p+0, 0x1.5ab07dd485429p+0};
  const uint32_t TBLBITS = 4, TBLSIZE = 16;
  const float redux = 0x1.8p23f / 16.0f;
  const float P1 = 0x1.62e430p-1f, P2 = 0x1.ebfbe0p-3f, P3 = 0x1.c6b348p-5f, P4 = 0x1.3b2c9cp-7f;
  volatile float t = x + redux;
  float tt = t;
  uint32_t i0;
  memcpy(&i0, &tt, 4);
  i0 += TBLSIZE / 2;
  uint32_t k = (i0 >> TBLBITS) << 20;
  i0 &= TBLSIZE - 1;
  tt -= redux;
  float z = x - tt;
  uint64_t bits = ((uint64_t)(0x3ff00000u + k)) << 32;
  double twopk_d;
  memcpy(&twopk_d, &bits, 8);
  float twopk = (float)twopk_d;
  float tv = (float)tbl[i0];
  float u = tv * z;
  tv = tv + u * (P1 + z * P2) + u * (z * z) * (P3 + z * P4);
  return tv * twopk;
}
static double fastpow(double x, double y) {
  if (x == 0.0) return 0.0;
  return (double)exp2f_tbl((float)y * fastlog2f((float)x));
}

/* Tsit5 tableau (Tsitouras 2011), as used by OrdinaryDiffEq's Tsit5ConstantCache */
static const double c1 = 0.161, c2 = 0.327, c3 = 0.9, c4 = 0.9800255409045097;
static const double a21 = 0.161, a31 = -0.008480655492356989, a32 = 0.335480655492357,
                    a41 = 2.8971530571054935, a42 = -6.359448489975075, a43 = 4.3622954328695815,
                    a51 = 5.325864828439257, a52 = -11.748883564062828, a53 = 7.4955393428898365,
                    a54 = -0.09249506636175525, a61 = 5.86145544294642, a62 = -12.92096931784711,
                    a63 = 8.159367898576159, a64 = -0.071584973281401, a65 = -0.028269050394068383,
                    a71 = 0.09646076681806523, a72 = 0.01, a73 = 0.4798896504144996,
                    a74 = 1.379008574103742, a75 = -3.290069515436081, a76 = 2.324710524099774;
static const double bt1 = -0.00178001105222577714, bt2 = -0.0008164344596567469,
                    bt3 = 0.007880878010261995, bt4 = -0.1447110071732629, bt5 = 0.5823571654525552,
                    bt6 = -0.45808210592918697, bt7 = 0.015151515151515152;

/* positions of the N+9 evolving entries inside the full state */
static void evolving_index(const sidx_t* ix, int N, int* pos) {
  for (int i = 0; i < N; i++) pos[i] = ix->SWATI + i;
  pos[N] = ix->GWAT;
  pos[N + 1] = ix->ACC + LWFB_A_FLOW;
  pos[N + 2] = ix->ACC + LWFB_A_SEEP;
  pos[N + 3] = ix->ACC + LWFB_A_SRFL;
  pos[N + 4] = ix->ACC + LWFB_A_SLFL;
  pos[N + 5] = ix->ACC + LWFB_A_BYFL;
  pos[N + 6] = ix->ACC + LWFB_A_DSFL;
  pos[N + 7] = ix->ACC + LWFB_A_GWFL;
  pos[N + 8] = ix->ACC + LWFB_A_VRFLN;
}

static double eps_of(double x) { /* Julia eps(x) */
  x = fabs(x);
  return nextafter(x, INFINITY) - x;
}

/*
 * Integrates ONE member over tspan = (t0_day, t0_day + n_days).
 *   out_daily [n_days+1][L]  saved records (saveat = t0:1:t1, saved BEFORE the callbacks of that
 *                            instant)
 *   counters  [3]            naccept, nreject, nf
 *   ntfli_init               NULL -> first ITER sees NaN (fresh problem); else warm caches
 *                            
 * returns the member status (LWFB_ST_*).
 */
int lwfb_oracle_run_member(int N, int iso, int t0_day, int n_days, int n_rows, const double* forcing,
                           const double* scal, const double* lay, const double* u0, double reltol,
                           double abstol, double dt_init, double dtmax, int64_t maxiters,
                           const double* ntfli_init, double* out_daily, int64_t* counters,
                           double* ntfli_final) {
  if (N < 1 || N > MAXN) return LWFB_ST_INPUT;
  model_t* m = (model_t*)calloc(1, sizeof(model_t));
  m->N = N; m->iso = iso; m->scal = scal; m->lay = lay; m->forc = forcing; m->n_rows = n_rows;
  m->RNET = m->SMLT = m->SLVP = m->TA = NAN;
  for (int i = 0; i < N; i++) {
    m->TRANI[i] = NAN;
    m->NTFLI[i] = ntfli_init ? ntfli_init[i] : NAN;
    m->VRFLI[i] = m->INFLI[i] = m->BYFLI[i] = m->DSFLI[i] = NAN;
  }
  sidx_t ix = make_idx(N, iso);
  int L = ix.L, NE = N + 9;
  int pos[MAXN + 9];
  evolving_index(&ix, N, pos);
  double* u = (double*)malloc(sizeof(double) * L);
  double* uprev = (double*)malloc(sizeof(double) * L);
  double* tmp = (double*)malloc(sizeof(double) * L);
  double k[7][MAXN + 9];
  memcpy(u, u0, sizeof(double) * L);
  const double qmin = 0.2, qmax = 10.0, gamma = 0.9, beta1 = 0.14, beta2 = 0.08, qoldinit = 1e-4;
  double t0 = (double)t0_day, tend = (double)(t0_day + n_days);
  double t = t0, tprev = t0, dt = dt_init, dtpropose = dt_init, qold = qoldinit, q11 = 1.0;
  double dtmin_opt = eps_of(tend) > eps_of(t0) ? eps_of(tend) : eps_of(t0);
  int64_t naccept = 0, nreject = 0, iter = 0;
  int accept = 0, reeval = 0;
  int status = LWFB_ST_OK;
  /* --- init: save u0, initial_affect of the daily callbacks, FSAL --- */
  memcpy(out_daily, u, sizeof(double) * L);
  cb_amounts(m, u, &ix, t0_day);
  cb_deltas(m, u, &ix, t0);
  memcpy(uprev, u, sizeof(double) * L);
  rhs(m, uprev + ix.SWATI, uprev[ix.GWAT], t, k[0]);
  double tstop = t0 + 1.0;
  if (tstop > tend) tstop = tend;
  int rec = 1;
  status = m->status;   /* a failing initial callback (RSS error) stops the member before its first step */
  while (status == LWFB_ST_OK && t < tend) {
    /* loopheader! */
    if (iter > 0) {
      if (accept) {
        memcpy(uprev, u, sizeof(double) * L);
        dt = dtpropose;
        if (reeval) rhs(m, u + ix.SWATI, u[ix.GWAT], t, k[0]);
        else memcpy(k[0], k[6], sizeof(double) * NE);
      } else {
        dt = dt / jmin(1 / qmin, q11 / gamma);
      }
    }
    iter++;
    if (iter > maxiters) { status = LWFB_ST_MAXITERS; break; }
    {
      int bad = 0;
      for (int i = 0; i < N; i++) if (isnan(u[ix.SWATI + i])) bad = 1;
      if (bad) { status = LWFB_ST_UNSTABLE; break; }
    }
    dt = dt < dtmax ? dt : dtmax;
    {
      double dtmin_t = eps_of(t) > dtmin_opt ? eps_of(t) : dtmin_opt;
      if (dt < dtmin_t) dt = dtmin_t;
    }
    if (dt > tstop - t) dt = tstop - t;
    if (!(fabs(dt) > dtmin_opt)) { status = LWFB_ST_DTMIN; break; }
    /* perform_step! (Tsit5Cache, in place) */
    memcpy(tmp, uprev, sizeof(double) * L);
    for (int j = 0; j < NE; j++) tmp[pos[j]] = uprev[pos[j]] + (dt * a21) * k[0][j];
    rhs(m, tmp + ix.SWATI, tmp[ix.GWAT], t + c1 * dt, k[1]);
    for (int j = 0; j < NE; j++) tmp[pos[j]] = uprev[pos[j]] + dt * (a31 * k[0][j] + a32 * k[1][j]);
    rhs(m, tmp + ix.SWATI, tmp[ix.GWAT], t + c2 * dt, k[2]);
    for (int j = 0; j < NE; j++) tmp[pos[j]] = uprev[pos[j]] + dt * (a41 * k[0][j] + a42 * k[1][j] + a43 * k[2][j]);
    rhs(m, tmp + ix.SWATI, tmp[ix.GWAT], t + c3 * dt, k[3]);
    for (int j = 0; j < NE; j++)
      tmp[pos[j]] = uprev[pos[j]] + dt * (a51 * k[0][j] + a52 * k[1][j] + a53 * k[2][j] + a54 * k[3][j]);
    rhs(m, tmp + ix.SWATI, tmp[ix.GWAT], t + c4 * dt, k[4]);
    for (int j = 0; j < NE; j++)
      tmp[pos[j]] = uprev[pos[j]] + dt * (a61 * k[0][j] + a62 * k[1][j] + a63 * k[2][j] + a64 * k[3][j] + a65 * k[4][j]);
    rhs(m, tmp + ix.SWATI, tmp[ix.GWAT], t + dt, k[5]);
    for (int j = 0; j < NE; j++)
      u[pos[j]] = uprev[pos[j]] + dt * (a71 * k[0][j] + a72 * k[1][j] + a73 * k[2][j] + a74 * k[3][j] + a75 * k[4][j] + a76 * k[5][j]);
    rhs(m, u + ix.SWATI, u[ix.GWAT], t + dt, k[6]);   /* clamps u itself */
    /* residuals + norm_to_use (src/func_DiffEq_definition_ode.jl:20-37): RMS over 3N+39 entries,
     * of which only the N+9 evolving ones are non-zero */
    double ss = 0;
    for (int j = 0; j < NE; j++) {
      double ut = dt * (bt1 * k[0][j] + bt2 * k[1][j] + bt3 * k[2][j] + bt4 * k[3][j] + bt5 * k[4][j] + bt6 * k[5][j] + bt7 * k[6][j]);
      double r = ut / (abstol + jmax(fabs(uprev[pos[j]]), fabs(u[pos[j]])) * reltol);
      ss += r * r;
    }
    double EEst = sqrt(ss / (double)(3 * N + 39));
    /* loopfooter!: PI controller */
    reeval = 0;
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
      if (1.0 <= q && q <= 1.0) q = 1.0;
      qold = jmax(EEst, qoldinit);
      double dtnew = dt / q;
      tprev = t;
      {
        double mx = t > tstop ? t : tstop;
        if (fabs(ttmp - tstop) < 100 * eps_of(mx)) t = tstop; else t = ttmp;
      }
      dtpropose = dtmax < dtnew ? dtmax : dtnew;
      {
        double dtmin_t = eps_of(t) > dtmin_opt ? eps_of(t) : dtmin_opt;
        if (dtpropose < dtmin_t) dtpropose = dtmin_t;
      }
      /* handle_callbacks!: the first firing callback saves before its affect */
      if (t == tstop) {
        memcpy(out_daily + (size_t)rec * L, u, sizeof(double) * L);
        rec++;
        if (t < tend) {
          int day = (int)floor(t + 0.5);
          cb_balance(m, u, uprev, &ix);
          cb_ade(m, u, uprev, &ix, t, tprev);
          cb_amounts(m, u, &ix, day);
          cb_deltas(m, u, &ix, t);
          reeval = 1;
          tstop = tstop + 1.0;
          if (tstop > tend) tstop = tend;
        }
      } else {
        cb_ade(m, u, uprev, &ix, t, tprev);
      }
    } else {
      nreject++;
    }
    if (m->status != LWFB_ST_OK) { status = m->status; break; }
  }
  /* records not reached (aborted member) are filled with NaN */
  for (; rec <= n_days; rec++)
    for (int j = 0; j < L; j++) out_daily[(size_t)rec * L + j] = NAN;
  counters[LWFB_C_NACCEPT] = naccept;
  counters[LWFB_C_NREJECT] = nreject;
  counters[LWFB_C_NF] = m->nf;
  if (ntfli_final) memcpy(ntfli_final, m->NTFLI, sizeof(double) * N);
  free(u); free(uprev); free(tmp); free(m);
  return status;
}

/* Batch driver: members are independent; n_threads POSIX threads pull members from a shared
 * counter (the reference itself is single-threaded). */
#include <pthread.h>
typedef struct {
  int N, iso, t0_day, n_days, n_rows;
  const double* forcing;
  int64_t n_members;
  const int32_t* site_of_member;
  const double *scal, *lay, *u0;
  double reltol, abstol, dt_init, dtmax;
  int64_t maxiters;
  double* out_daily;
  int32_t* status;
  int64_t* counters;
  int64_t next;
  pthread_mutex_t mu;
} batch_t;

static void* batch_worker(void* arg) {
  batch_t* b = (batch_t*)arg;
  int L = LWFB_STATE_LEN(b->N, b->iso);
  double* scratch = b->out_daily ? NULL : (double*)malloc(sizeof(double) * (size_t)(b->n_days + 1) * L);
  for (;;) {
    pthread_mutex_lock(&b->mu);
    int64_t mi = b->next++;
    pthread_mutex_unlock(&b->mu);
    if (mi >= b->n_members) break;
    const double* f = b->forcing + (size_t)b->site_of_member[mi] * b->n_rows * LWFB_F;
    b->status[mi] = lwfb_oracle_run_member(
        b->N, b->iso, b->t0_day, b->n_days, b->n_rows, f, b->scal + (size_t)mi * LWFB_PS,
        b->lay + (size_t)mi * LWFB_PL * b->N, b->u0 + (size_t)mi * L, b->reltol, b->abstol, b->dt_init,
        b->dtmax, b->maxiters, NULL,
        b->out_daily ? b->out_daily + (size_t)mi * (b->n_days + 1) * L : scratch,
        b->counters + (size_t)mi * LWFB_NCOUNTERS, NULL);
  }
  free(scratch);
  return NULL;
}

int lwfb_oracle_run(int N, int iso, int t0_day, int n_days, int n_sites, int n_rows, const double* forcing,
                    int64_t n_members, const int32_t* site_of_member, const double* scal, const double* lay,
                    const double* u0, double reltol, double abstol, double dt_init, double dtmax,
                    int64_t maxiters, int n_threads, double* out_daily, int32_t* status, int64_t* counters) {
  (void)n_sites;
  batch_t b;
  b.N = N; b.iso = iso; b.t0_day = t0_day; b.n_days = n_days; b.n_rows = n_rows; b.forcing = forcing;
  b.n_members = n_members; b.site_of_member = site_of_member; b.scal = scal; b.lay = lay; b.u0 = u0;
  b.reltol = reltol; b.abstol = abstol; b.dt_init = dt_init; b.dtmax = dtmax; b.maxiters = maxiters;
  b.out_daily = out_daily; b.status = status; b.counters = counters; b.next = 0;
  pthread_mutex_init(&b.mu, NULL);
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  for (int i = 1; i < n_threads; i++) pthread_create(&th[i], NULL, batch_worker, &b);
  batch_worker(&b);
  for (int i = 1; i < n_threads; i++) pthread_join(th[i], NULL);
  pthread_mutex_destroy(&b.mu);
  return 0;
}

/* single RHS evaluation, for unit parity tests against lwfb_eval_rhs */
void lwfb_oracle_eval_rhs(int N, const double* scal, const double* lay, double t, double* swati, double gwat,
                          const double* frozen, const double* ntfli_prev, double* du) {
  model_t* m = (model_t*)calloc(1, sizeof(model_t));
  m->N = N; m->scal = scal; m->lay = lay;
  m->RNET = frozen[0]; m->SMLT = frozen[1]; m->SLVP = frozen[2];
  for (int i = 0; i < N; i++) { m->TRANI[i] = frozen[3 + i]; m->NTFLI[i] = ntfli_prev[i]; }
  rhs(m, swati, gwat, t, du);
  free(m);
}

/* exposed leaf functions for known-answer tests (test/01-unit-tests.jl:41-66) */
double lwfb_oracle_kkmean(double K1, double K2, double T1, double T2) {
  return exp((T2 * log(K1) + T1 * log(K2)) / (T1 + T2));
}
void lwfb_oracle_esat(double TA, double* ES, double* DELTA) { esat(TA, ES, DELTA); }
double lwfb_oracle_pm(double AA, double VPD, double DELTA, double RA, double RC) { return pm(AA, VPD, DELTA, RA, RC); }
double lwfb_oracle_fastpow(double x, double y) { return fastpow(x, y); }
/* elementary functions as compiled into this build (libm in mode A, lwfb90_elem.h in mode B) */
double lwfb_oracle_elem(int fn, double x, double y) {
  switch (fn) {
    case 0: return exp(x);
    case 1: return log(x);
    case 2: return pow(x, y);
    case 3: return sin(x);
    case 4: return cos(x);
    case 5: return tan(x);
    case 6: return asin(x);
    case 7: return acos(x);
    default: return NAN;
  }
}
