/*
 * lwfb90_setup_twin.c — CPU twin of the device-side batched setup.  TEST INFRASTRUCTURE ONLY (see oracle.py).
 *
 * Compiles include/lwfb90_setup.h (the text the CUDA kernel setup_members_kernel runs, with the elementary functions of
 * include/lwfb90_elem.h) for the host, so that tests can require the device-built member blocks to be BIT-IDENTICAL to
 * the host-built ones; the independent check is lwfbrook90.jl_b200/ensemble.py:batched_setup (numpy, host libm), pinned
 * to the reference's remakeSPAC -> setup (src/LWFBrook90.jl:214-345, :365-587) by tests/test_host.py.
 */
#include "../include/lwfb90_setup.h"

void lwfb_oracle_setup_members(int64_t M, int N, int iso, const double* base_scal, const double* base_lay,
                               const double* base_u0, const double* knobs, int n_over, const int32_t* over_idx,
                               int n_profiles, const double* profiles, const double* psim_base, int ilayer, double* scal,
                               double* lay, double* u0, int32_t* status) {
  const int L = LWFB_STATE_LEN(N, iso);
  for (int64_t m = 0; m < M; m++)
    status[m] = lwfb_setup_member(N, iso, base_scal, base_lay, base_u0, knobs + m * (LWFB_K + n_over), n_over, over_idx,
                                  n_profiles, profiles, psim_base, ilayer, scal + m * LWFB_PS, lay + m * (int64_t)LWFB_PL * N,
                                  u0 + m * L);
}
