// lwfb90_setup.cuh — batched `setup` on the device: one thread per member builds the member's
// parameter blocks and initial state in HBM from the base member and the member's knob row (include/lwfb90_setup.h has
// the arithmetic and the reference citations).  The blocks are member-major, so every thread writes its own contiguous
// piece; the work per member is ~54 bisection steps x N layers x 3 powers, i.e. microseconds for the whole ensemble.
#pragma once
#include "../../include/lwfb90_setup.h"

namespace lwfb {

__global__ void setup_members_kernel(int64_t M, int N, int iso, const double* __restrict__ base_scal,
                                     const double* __restrict__ base_lay, const double* __restrict__ base_u0,
                                     const double* __restrict__ knobs, int n_over, const int32_t* __restrict__ over_idx,
                                     int n_profiles, const double* __restrict__ profiles,
                                     const double* __restrict__ psim_base, int ilayer, double* __restrict__ scal,
                                     double* __restrict__ lay, double* __restrict__ u0, int32_t* __restrict__ status) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int L = LWFB_STATE_LEN(N, iso);
  status[m] = lwfb_setup_member(N, iso, base_scal, base_lay, base_u0, knobs + m * (LWFB_K + n_over), n_over, over_idx,
                                n_profiles, profiles, psim_base, ilayer, scal + m * LWFB_PS, lay + m * (int64_t)LWFB_PL * N,
                                u0 + m * L);
}

}  // namespace lwfb
