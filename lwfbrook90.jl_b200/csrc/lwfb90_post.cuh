// lwfb90_post.cuh — post-processing of the daily record on the device.
//
// water_partitioning_kernel: the ensemble form of get_water_partitioning (src/func_postprocess.jl:1310-1364):
// per member and per period (month, year, ... any grouping of the daily records) the sums of
//   ETa, Esoil, Esnow, Einterception, Ta, Precip, Td, D, R     and the mean of Swat
// from the accumulator entries of the daily record, so that a 36 GB record never has to leave the GPU when only the
// partitioning is wanted.  HBM-bound.  The record is member-major ([M][n_days+1][L]): one WARP per member, lane e reads
// accumulator entry e of a record (the 31 entries are one contiguous 248-byte piece), the ten columns are formed in
// lanes 0..9 from shuffled entries and summed sequentially in record order (the order of a cumulative sum on the host).
#pragma once
#include "lwfb90_device.cuh"

namespace lwfb {

// out [M][n_groups][LWFB_NWP]; group_of_record [n_days+1]: period index of each daily record, < 0 = not used.
// Records of one period are expected to be contiguous (months, years); a period that comes back after another one
// continues its sums (partial sums added in order of appearance).
__global__ void water_partitioning_kernel(const double* __restrict__ rec, int64_t M, int L, int acc0, int n_rec,
                                          const int32_t* __restrict__ group_of_record, int n_groups,
                                          double* __restrict__ out) {
  const int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= M) return;                                  // (whole warps leave together)
  // column `lane` = kind(x1, x2, x3) of the accumulator entries e1, e2, e3, written as on the host
  int e1 = 0, e2 = 0, e3 = 0, kind = 0;                // kind 0: x1; 1: x1 + x2; 2: x1 - x2; 3: -x1; 4: -((x1 + x2) + x3)
  switch (lane) {
    case LWFB_WP_ETA: e1 = LWFB_A_CUM_D_EVAP; break;
    case LWFB_WP_ESOIL: e1 = LWFB_A_CUM_D_SLVP; break;
    case LWFB_WP_ESNOW: e1 = LWFB_A_CUM_D_SNVP; break;
    case LWFB_WP_EINTERCEPTION: e1 = LWFB_A_CUM_D_IRVP; e2 = LWFB_A_CUM_D_ISVP; kind = 1; break;
    case LWFB_WP_TA: e1 = LWFB_A_CUM_D_TRAN; break;
    case LWFB_WP_PRECIP: e1 = LWFB_A_CUM_D_PREC; break;
    case LWFB_WP_TD: e1 = LWFB_A_CUM_D_PTRAN; e2 = LWFB_A_CUM_D_TRAN; kind = 2; break;
    case LWFB_WP_D: e1 = LWFB_A_VRFLN; kind = 3; break;
    case LWFB_WP_R: e1 = LWFB_A_SRFL; e2 = LWFB_A_BYFL; e3 = LWFB_A_DSFL; kind = 4; break;
    case LWFB_WP_SWAT: e1 = LWFB_A_STORAGESWAT; break;
    default: break;
  }
  const bool col = lane < LWFB_NWP;
  double* o = out + m * (int64_t)n_groups * LWFB_NWP;
  for (int k = lane; k < n_groups * LWFB_NWP; k += 32) o[k] = 0.0;
  __syncwarp();
  const double* rm = rec + m * (int64_t)n_rec * L + acc0;
  double s = 0.0;
  int g_cur = -1;
  for (int r = 0; r < n_rec; r++) {
    const int g = __ldg(group_of_record + r);         // uniform over the grid
    if (g != g_cur) {
      if (g_cur >= 0 && col) o[(int64_t)g_cur * LWFB_NWP + lane] += s;
      s = 0.0;
      g_cur = (g >= 0 && g < n_groups) ? g : -1;
    }
    if (g_cur < 0) continue;
    const double a = lane < LWFB_NACCUM ? rm[(int64_t)r * L + lane] : 0.0;
    const double x1 = __shfl_sync(0xffffffffu, a, e1), x2 = __shfl_sync(0xffffffffu, a, e2), x3 = __shfl_sync(0xffffffffu, a, e3);
    double v = x1;
    if (kind == 1) v = x1 + x2;
    else if (kind == 2) v = x1 - x2;
    else if (kind == 3) v = -x1;
    else if (kind == 4) v = -((x1 + x2) + x3);
    s += v;
  }
  if (g_cur >= 0 && col) o[(int64_t)g_cur * LWFB_NWP + lane] += s;
}

// Swat is a mean: divide its sums by the number of records of the period
__global__ void water_partitioning_mean(double* __restrict__ out, int64_t M, int n_groups,
                                        const int32_t* __restrict__ n_in_group) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= M * n_groups) return;
  const int g = (int)(k % n_groups);
  const int n = n_in_group[g];
  double* v = out + k * LWFB_NWP + LWFB_WP_SWAT;
  *v = n > 0 ? *v / (double)n : NAN;
}

// soil_state_kernel: the ensemble form of get_soil_ (src/func_postprocess.jl:1025-1045 with the soil variables of
// derive_auxiliary_SOILVAR, src/module_KPT.jl:432-443): for every member, selected daily records and selected layers
//   theta (-), psi (kPa), W (-), SWATI (mm), K (mm/day), delta18O, delta2H (permil; NaN without isotopes)
// out [M][n_rec_sel][n_lay_sel][LWFB_NSV].  Same expressions as the daily callback of the integrator.
__global__ void soil_state_kernel(const double* __restrict__ rec, const double* __restrict__ lay, int64_t M, int N, int L,
                                  int iso, int n_rec, int n_rec_sel, const int32_t* __restrict__ rec_idx, int n_lay_sel,
                                  const int32_t* __restrict__ lay_idx, double* __restrict__ out) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  auto LP = [&](int field, int i) { return __ldg(lay + (m * LWFB_PL + field) * N + i); };
  for (int q = 0; q < n_lay_sel; q++) {
    const int i = lay_idx[q];
    const double THSAT = LP(LWFB_L_THSAT, i), THR = LP(LWFB_L_THR, i), SWATMAX = LP(LWFB_L_SWATMAX, i);
    const double NPAR = LP(LWFB_L_NPAR, i), MPAR = LP(LWFB_L_MPAR, i);
    const double DTH = THSAT - THR, E1 = -1.0 / MPAR, E2 = 1.0 / NPAR, E3 = NPAR / (NPAR - 1.0);
    const double CPSI = 9.81 * (-1.0 / LP(LWFB_L_ALPHA, i)), KSAT = LP(LWFB_L_KSAT, i), TORT = LP(LWFB_L_TORT, i);
    for (int k = 0; k < n_rec_sel; k++) {
      const double* r = rec + (m * (int64_t)n_rec + rec_idx[k]) * L;
      const double sw = r[i];
      const double w = jminc((THSAT * sw / SWATMAX - THR) / DTH, 1.0);
      const double a = jmaxc(w, 1.e-6);
      const double la = lwfb_log(a);
      const double p1 = lwfb_pow_from_log(a, la, E1);
      const double psim = CPSI * lwfb_pow(p1 - 1.0, E2);
      const double p3 = lwfb_pow_from_log(a, la, E3);
      const double inner = 1.0 - lwfb_pow(1.0 - p3, 1.0 - E2);
      const double kk = KSAT * lwfb_pow_from_log(a, la, TORT) * (inner * inner);
      double* o = out + ((m * n_rec_sel + k) * (int64_t)n_lay_sel + q) * LWFB_NSV;
      o[LWFB_SV_THETA] = w * DTH + THR;
      o[LWFB_SV_PSI] = psim;
      o[LWFB_SV_W] = w;
      o[LWFB_SV_SWATI] = sw;
      o[LWFB_SV_K] = kk;
      o[LWFB_SV_D18O] = iso ? r[N + i] : NAN;
      o[LWFB_SV_D2H] = iso ? r[2 * N + i] : NAN;
    }
  }
}

}  // namespace lwfb
