// lwfb90_member.cu — thread-per-member kernels: the fallback integrator for layer counts outside the
// lane-per-layer kernel's range (N < 3 or N > 32) and the single-RHS kernel of the unit parity tests.
//
// Own translation unit so that it compiles in parallel with lwfb90.cu.  (History: an earlier version of the daily
// callback selected between the addresses of two thread-local arrays inside a non-unrolled loop,
// `double* ATRI = (J == 1) ? ATRI0 : ATRI1;`; with NVVM -O2/-O3 (nvcc 12.9, sm_100a) this kernel then produced results
// that differed from the CPU oracle in the 4th digit, while -O0, -O1, -G and a host build of the same text agreed
// with it bit for bit.  One two-dimensional array instead of the pointer select removed the problem,
// profiles/r01_notes.md.)  These kernels are not on the measured path.
#include "lwfb90_launch.h"
#include "lwfb90_member.cuh"

using namespace lwfb;

#ifndef LWFB_BLOCK
#define LWFB_BLOCK 128
#endif

template <int NT>
__global__ void __launch_bounds__(LWFB_BLOCK) integrate_kernel(DevData d) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= d.M) return;
  integrate_member<NT>(d, m);
}

// one RHS evaluation per member (parity tests)
template <int NT>
__global__ void rhs_kernel(DevData d, double t, const double* __restrict__ swati, const double* __restrict__ gwat,
                           const double* __restrict__ frozen, const double* __restrict__ ntfli_prev,
                           double* __restrict__ du) {
  constexpr int NMAX = Member<NT>::NMAX;
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= d.M) return;
  Member<NT> mb;
  mb.load(d, m);
  const int N = mb.N;
  double sw[NMAX], out[NMAX + 9];
  for (int i = 0; i < N; i++) sw[i] = swati[m * N + i];
  mb.RNET = frozen[m * (N + 3) + 0]; mb.SMLT = frozen[m * (N + 3) + 1]; mb.SLVP = frozen[m * (N + 3) + 2];
  for (int i = 0; i < N; i++) { mb.TRANI[i] = frozen[m * (N + 3) + 3 + i]; mb.NTFLI[i] = ntfli_prev[m * N + i]; }
  mb.rhs(sw, gwat[m], t, out);
  for (int j = 0; j < N + 9; j++) du[m * (N + 9) + j] = out[j];
}


namespace lwfb {
void launch_member_integrate(const DevData& d, cudaStream_t st) {
  unsigned grid = (unsigned)((d.M + LWFB_BLOCK - 1) / LWFB_BLOCK);
  integrate_kernel<0><<<grid, LWFB_BLOCK, 0, st>>>(d);
}
void launch_member_rhs(const DevData& d, double t, const double* swati, const double* gwat, const double* frozen,
                       const double* ntfli_prev, double* du) {
  unsigned grid = (unsigned)((d.M + 127) / 128);
  rhs_kernel<0><<<grid, 128>>>(d, t, swati, gwat, frozen, ntfli_prev, du);
}
}  // namespace lwfb
