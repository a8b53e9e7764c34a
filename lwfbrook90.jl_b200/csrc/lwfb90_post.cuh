// lwfb90_post.cuh — post-processing of the daily record on the device.
//
// water_partitioning_kernel: the ensemble form of get_water_partitioning (src/func_postprocess.jl:1310-1364):
// per member and per period (month, year, ... any grouping of the daily records) the sums of
//   ETa, Esoil, Esnow, Einterception, Ta, Precip, Td, D, R     and the mean of Swat
// from the accumulator entries of the daily record, so that a 36 GB record never has to leave the GPU when only the
// partitioning is wanted.  HBM-bound: 12 doubles read per member and record (coalesced over the members), one thread
// per member, sequential summation in record order (the order of a cumulative sum on the host).
#pragma once
#include "lwfb90_device.cuh"

namespace lwfb {

// out [M][n_groups][LWFB_NWP]; group_of_record [n_days+1]: period index of each daily record, < 0 = not used.
// Records of one period are expected to be contiguous (months, years); a period that comes back after another one
// continues its sums (partial sums added in order of appearance).
__global__ void water_partitioning_kernel(const double* __restrict__ rec, int64_t M, int L, int acc0, int n_rec,
                                          const int32_t* __restrict__ group_of_record, int n_groups,
                                          double* __restrict__ out) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double* o = out + m * (int64_t)n_groups * LWFB_NWP;
  for (int k = 0; k < n_groups * LWFB_NWP; k++) o[k] = 0.0;
  double s[LWFB_NWP];
#pragma unroll
  for (int c = 0; c < LWFB_NWP; c++) s[c] = 0.0;
  int g_cur = -1;
  auto flush = [&]() {
    if (g_cur < 0) return;
    double* og = o + (int64_t)g_cur * LWFB_NWP;
#pragma unroll
    for (int c = 0; c < LWFB_NWP; c++) og[c] += s[c];
#pragma unroll
    for (int c = 0; c < LWFB_NWP; c++) s[c] = 0.0;
  };
  for (int r = 0; r < n_rec; r++) {
    const int g = __ldg(group_of_record + r);         // uniform over the grid
    if (g != g_cur) { flush(); g_cur = (g >= 0 && g < n_groups) ? g : -1; }
    if (g_cur < 0) continue;
    const double* a = rec + ((int64_t)r * L + acc0) * M + m;
    auto A = [&](int e) { return a[(int64_t)e * M]; };
    const double tran = A(LWFB_A_CUM_D_TRAN);
    s[LWFB_WP_ETA] += A(LWFB_A_CUM_D_EVAP);
    s[LWFB_WP_ESOIL] += A(LWFB_A_CUM_D_SLVP);
    s[LWFB_WP_ESNOW] += A(LWFB_A_CUM_D_SNVP);
    s[LWFB_WP_EINTERCEPTION] += A(LWFB_A_CUM_D_IRVP) + A(LWFB_A_CUM_D_ISVP);
    s[LWFB_WP_TA] += tran;
    s[LWFB_WP_PRECIP] += A(LWFB_A_CUM_D_PREC);
    s[LWFB_WP_TD] += A(LWFB_A_CUM_D_PTRAN) - tran;
    s[LWFB_WP_D] += -A(LWFB_A_VRFLN);
    s[LWFB_WP_R] += -(A(LWFB_A_SRFL) + A(LWFB_A_BYFL) + A(LWFB_A_DSFL));
    s[LWFB_WP_SWAT] += A(LWFB_A_STORAGESWAT);
  }
  flush();
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

}  // namespace lwfb
