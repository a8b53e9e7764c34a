/*
 * lwfb90_setup.h — batched `setup` of ensemble members on the device, in portable C.
 *
 * The reference builds one member at a time: remakeSPAC -> setup (src/LWFBrook90.jl:214-345, :365-587) runs
 * define_LWFB90_p (src/func_DiffEq_definition_p.jl:12-586), the KPT_SOILPAR_Mvg1d constructor with its bisection for
 * the wetness at field capacity (src/module_KPT.jl:333-406) and init_LWFB90_u0! (src/func_DiffEq_definition_u0.jl:84-139).
 * lwfb_setup_member() below does the part of that which depends on the knobs remakeSPAC can turn, for ONE
 * member, from the flattened blocks of the unmodified site definition ("base member") and the member's knob values:
 *
 *   soil horizons   ths, Ksat, alpha, (n - 1), gravel fraction scaled by per-member factors of the top-horizon value,
 *                   all horizons proportionally (remake_soil_horizons, src/LWFBrook90.jl:266-272)
 *                   -> SWATMAX, WETF (bisection of FK_MvG(W) = 2 mm/d), PSIF, THETAF, SWATQX, SWATQF
 *   params          any LWFB_S_* scalar overwritten; INFEXP -> INFRAC (INFPAR, src/module_WAT.jl:59-79)
 *   roots           one of a small table of root profiles (beta model quantised to a few levels)
 *   initial state   uniform initial matric potential -> SWATI, StorageSWAT, StorageWATER
 *
 * The same text compiles for the device (one thread per member, csrc/lwfb90_setup.cuh) and for a host C compiler (the
 * CPU twin in oracle/, test infrastructure): both evaluate the powers with include/lwfb90_elem.h and give bit-identical
 * blocks.  The numpy implementation lwfbrook90.jl_b200/ensemble.py:batched_setup (host libm) agrees to ~1e-13.
 */
#ifndef LWFB90_SETUP_H
#define LWFB90_SETUP_H

#include "lwfb90.h"
#include "lwfb90_elem.h"

/* per-member knob columns of the `knobs` block [n_members][LWFB_K]; NaN = leave the base value */
enum {
  LWFB_K_THS_F = 0,   /* factor on ths of every horizon (capped at 0.95)                   */
  LWFB_K_KSAT_F,      /* factor on Ksat                                                    */
  LWFB_K_ALPHA_F,     /* factor on alpha                                                   */
  LWFB_K_NM1_F,       /* factor on (n - 1) of the top horizon, the others proportionally   */
  LWFB_K_GRAVEL_F,    /* factor on the gravel fraction (capped at 0.95)                    */
  LWFB_K_INFEXP,      /* INFEXP (-> INFRAC)                                                */
  LWFB_K_ROOTPROFILE, /* row of the root-profile table                                     */
  LWFB_K_PSIM_INIT,   /* uniform initial matric potential, kPa                             */
  LWFB_K              /* followed by n_over columns: values of the overridden LWFB_S_* scalars */
};

/* FK_MvG, src/module_KPT.jl:495-500 */
LWFB_ELEM double lwfb_setup_fk(double W, double KSAT, double l, double n) {
  const double A = !(W <= 1.0e-6) ? W : 1.0e-6;
  const double inner = 1 - lwfb_pow(1 - lwfb_pow(A, n / (n - 1)), 1 - 1 / n);
  return KSAT * lwfb_pow(A, l) * (inner * inner);
}

/* root of FK_MvG(W) - Kfc on (0, 1): bisection down to adjacent floats, what Roots.find_zero(f, (0.0, 1.0),
 * Bisection()) delivers (src/module_KPT.jl:359-361); same midpoint sequence as kpt.py:_bisect_wetf.  NaN: no root. */
LWFB_ELEM double lwfb_setup_wetf(double Kfc, double KSAT, double l, double n) {
  double a = 0.0, b = 1.0;
  double fa = -Kfc + lwfb_setup_fk(a, KSAT, l, n), fb = -Kfc + lwfb_setup_fk(b, KSAT, l, n);
  if (!(fa < 0 && fb > 0)) return NAN;
  for (int it = 0; it < 200; it++) {
    const double mid = 0.5 * (a + b);
    if (!(mid > a && mid < b)) break;
    const double fm = -Kfc + lwfb_setup_fk(mid, KSAT, l, n);
    if (fm == 0) { a = mid; b = mid; fa = fm; fb = fm; break; }
    if ((fm < 0) == (fa < 0)) { a = mid; fa = fm; } else { b = mid; fb = fm; }
  }
  return fabs(fa) <= fabs(fb) ? a : b;
}

/* One member.  base_* : blocks of the unmodified site definition (lwfb_set_members layouts for one member);
 * knob [LWFB_K + n_over]; over_idx [n_over] LWFB_S_* indices; profiles [n_profiles][2 N + 1] = RELDEN[N], TINI[N],
 * RELDEN_CONST; psim_base [N] initial matric potential of the soil discretisation (kPa); ilayer = number of layers
 * that receive infiltration (IDEPTH).  Outputs scal [LWFB_PS], lay [LWFB_PL][N], u0 [L].  Returns LWFB_ST_OK or
 * LWFB_ST_INPUT (no field-capacity wetness: the reference throws "Computed invalid p_WETF"). */
LWFB_ELEM int lwfb_setup_member(int N, int iso, const double* base_scal, const double* base_lay, const double* base_u0,
                                const double* knob, int n_over, const int32_t* over_idx, int n_profiles,
                                const double* profiles, const double* psim_base, int ilayer, double* scal, double* lay,
                                double* u0) {
  const int L = LWFB_STATE_LEN(N, iso);
  int status = LWFB_ST_OK;
  for (int k = 0; k < LWFB_PS; k++) scal[k] = base_scal[k];
  for (int k = 0; k < LWFB_PL * N; k++) lay[k] = base_lay[k];
  for (int k = 0; k < L; k++) u0[k] = base_u0[k];
#define LWFB_LY(f, i) lay[(f) * N + (i)]
  const double ths_f = knob[LWFB_K_THS_F], ksat_f = knob[LWFB_K_KSAT_F], alpha_f = knob[LWFB_K_ALPHA_F];
  const double nm1_f = knob[LWFB_K_NM1_F], gravel_f = knob[LWFB_K_GRAVEL_F];
  const int soil_changed = (ths_f == ths_f) || (ksat_f == ksat_f) || (alpha_f == alpha_f) || (nm1_f == nm1_f) || (gravel_f == gravel_f);
  double ratio = 1.0;
  if (nm1_f == nm1_f) {
    const double n_top = base_lay[LWFB_L_NPAR * N + 0];
    ratio = (1.0 + (n_top - 1.0) * nm1_f) / n_top;
  }
  for (int i = 0; i < N; i++) {
    if (ths_f == ths_f) { const double v = LWFB_LY(LWFB_L_THSAT, i) * ths_f; LWFB_LY(LWFB_L_THSAT, i) = v < 0.95 ? v : 0.95; }
    if (ksat_f == ksat_f) LWFB_LY(LWFB_L_KSAT, i) = LWFB_LY(LWFB_L_KSAT, i) * ksat_f;
    if (alpha_f == alpha_f) LWFB_LY(LWFB_L_ALPHA, i) = LWFB_LY(LWFB_L_ALPHA, i) * alpha_f;
    if (nm1_f == nm1_f) LWFB_LY(LWFB_L_NPAR, i) = LWFB_LY(LWFB_L_NPAR, i) * ratio;
    LWFB_LY(LWFB_L_MPAR, i) = 1.0 - 1.0 / LWFB_LY(LWFB_L_NPAR, i);
    if (gravel_f == gravel_f) { const double v = LWFB_LY(LWFB_L_STONEF, i) * gravel_f; LWFB_LY(LWFB_L_STONEF, i) = v < 0.95 ? v : 0.95; }
    LWFB_LY(LWFB_L_SWATMAX, i) = LWFB_LY(LWFB_L_THICK, i) * LWFB_LY(LWFB_L_THSAT, i) * (1.0 - LWFB_LY(LWFB_L_STONEF, i));
    if (soil_changed) {
      const double w = lwfb_setup_wetf(2.0, LWFB_LY(LWFB_L_KSAT, i), LWFB_LY(LWFB_L_TORT, i), LWFB_LY(LWFB_L_NPAR, i));
      if (w != w) status = LWFB_ST_INPUT;
      LWFB_LY(LWFB_L_WETF, i) = w;
      /* FPSIM_MvG :590-594, FTheta_MvG :683-685 */
      const double A = !(w <= 1.0e-6) ? w : 1.0e-6;
      LWFB_LY(LWFB_L_PSIF, i) = 9.81 * (-1 / LWFB_LY(LWFB_L_ALPHA, i)) *
                                lwfb_pow(lwfb_pow(A, -1 / LWFB_LY(LWFB_L_MPAR, i)) - 1, 1 / LWFB_LY(LWFB_L_NPAR, i));
      LWFB_LY(LWFB_L_THETAF, i) = w * (LWFB_LY(LWFB_L_THSAT, i) - LWFB_LY(LWFB_L_THR, i)) + LWFB_LY(LWFB_L_THR, i);
    }
  }
  for (int j = 0; j < n_over; j++) {
    const double v = knob[LWFB_K + j];
    if (v == v) scal[over_idx[j]] = v;
  }
  {
    const int Q = (int)base_scal[LWFB_S_QLAYER];
    double qx = 0, qf = 0;
    for (int i = 0; i < Q && i < N; i++) {
      qx = qx + LWFB_LY(LWFB_L_SWATMAX, i);
      qf = qf + LWFB_LY(LWFB_L_THETAF, i) * LWFB_LY(LWFB_L_THICK, i) * (1 - LWFB_LY(LWFB_L_STONEF, i));
    }
    scal[LWFB_S_SWATQX] = qx; scal[LWFB_S_SWATQF] = qf;
  }
  {
    const double e = knob[LWFB_K_INFEXP];
    if (e == e) {                                      /* WAT.INFPAR, src/module_WAT.jl:59-79 */
      for (int i = 0; i < N; i++) LWFB_LY(LWFB_L_INFRAC, i) = 0.0;
      if (e <= 0 || ilayer <= 0) LWFB_LY(LWFB_L_INFRAC, 0) = 1.0;
      else {
        double thickt = 0;
        for (int i = 0; i < ilayer && i < N; i++) thickt = thickt + LWFB_LY(LWFB_L_THICK, i);
        double thicka = 0, prev = lwfb_pow(0 / thickt, e);
        for (int i = 0; i < ilayer && i < N; i++) {
          thicka = thicka + LWFB_LY(LWFB_L_THICK, i);
          const double cur = lwfb_pow(thicka / thickt, e);
          LWFB_LY(LWFB_L_INFRAC, i) = cur - prev;
          prev = cur;
        }
      }
    }
  }
  {
    const double pr = knob[LWFB_K_ROOTPROFILE];
    if (pr == pr && n_profiles > 0) {
      int r = (int)pr;
      r = r < 0 ? 0 : (r >= n_profiles ? n_profiles - 1 : r);
      const double* p = profiles + (size_t)r * (2 * N + 1);
      for (int i = 0; i < N; i++) { LWFB_LY(LWFB_L_RELDEN, i) = p[i]; LWFB_LY(LWFB_L_TINI, i) = p[N + i]; }
      scal[LWFB_S_RELDEN_CONST] = p[2 * N];
    }
  }
  {
    /* init_LWFB90_u0!, src/func_DiffEq_definition_u0.jl:84-139: SWATI from the initial matric potential */
    const double pinit = knob[LWFB_K_PSIM_INIT];
    if (pinit == pinit || soil_changed) {
      const int k = iso ? 3 : 1;
      const int o_gwat = k * N, o_ints = o_gwat + k, o_intr = o_ints + k, o_snow = o_intr + k;
      const int o_acc = (iso ? 9 * N + 20 : 5 * N + 8);
      double new_swat = 0;
      for (int i = 0; i < N; i++) {
        const double psim0 = (pinit == pinit) ? pinit : psim_base[i];
        const double npar = LWFB_LY(LWFB_L_NPAR, i), mpar = LWFB_LY(LWFB_L_MPAR, i);
        const double W0 = (psim0 <= 0) ? lwfb_pow(1 + lwfb_pow(-LWFB_LY(LWFB_L_ALPHA, i) * psim0 / 9.81, npar), -mpar) : 1.0;   /* FWETNES :706-722 */
        const double th = W0 * (LWFB_LY(LWFB_L_THSAT, i) - LWFB_LY(LWFB_L_THR, i)) + LWFB_LY(LWFB_L_THR, i);
        const double sw0 = th * LWFB_LY(LWFB_L_SWATMAX, i) / LWFB_LY(LWFB_L_THSAT, i);
        u0[i] = sw0;
        new_swat = new_swat + sw0;
      }
      u0[o_acc + LWFB_A_STORAGESWAT] = new_swat;
      u0[o_acc + LWFB_A_STORAGEWATER] = u0[o_intr] + u0[o_ints] + u0[o_snow] + new_swat + u0[o_gwat];
    }
  }
#undef LWFB_LY
  return status;
}

#endif /* LWFB90_SETUP_H */
