// lwfb90_wide.cu — the warp-per-member integrator for deep profiles (N > 32), see lwfb90_wide.cuh.
//
// Own translation unit: it is compiled with -DLWFB_CB_SCRATCH, which moves the per-layer work arrays of the daily
// callback (cb_core, tbylayer in lwfb90_device.cuh) from thread-local arrays of a compile-time size into the member's
// shared-memory block, so that the layer count is only bounded by the shared memory of an SM.
#ifndef LWFB_CB_SCRATCH
#define LWFB_CB_SCRATCH 1
#endif
#include "lwfb90_launch.h"
#include "lwfb90_wide.cuh"

using namespace lwfb;

// one warp per member, `wpb` warps per block; shared memory: wpb blocks of LWFB_WIDE_DOUBLES_PER_MEMBER(N) doubles
__global__ void __launch_bounds__(256, 2) integrate_wide(DevData d, int wpb) {
  extern __shared__ double wide_smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * wpb + w;
  if (m >= d.M) return;                              // (whole warps leave; there are no block-level barriers)
  wide_body(d, lane, m, wide_smem + (size_t)w * LWFB_WIDE_DOUBLES_PER_MEMBER(d.N));
}

namespace lwfb {
int wide_max_layers(int device) {
  int optin = 0;
  if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return 0;
  // largest N whose member block fits the opt-in shared memory of one thread block
  const long long doubles = (long long)optin / (long long)sizeof(double);
  const long long n = doubles / WA_COUNT - 9;
  return (int)(n < 0 ? 0 : (n > LWFB_MAX_LAYERS ? LWFB_MAX_LAYERS : n));
}
cudaError_t launch_wide_integrate(const DevData& d, int device, cudaStream_t st) {
  int optin = 0, n_sm = 0;
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
  const size_t per_member = LWFB_WIDE_DOUBLES_PER_MEMBER(d.N) * sizeof(double);
  if (per_member > (size_t)optin) return cudaErrorInvalidConfiguration;
  // warps (members) per block: as many as fit half of the SM's shared memory (two resident blocks), at most 8; fewer
  // when the ensemble is small, so that the members spread over all SMs
  int wpb = (int)((size_t)optin / 2 / per_member);
  if (wpb < 1) wpb = 1;
  if (wpb > 8) wpb = 8;
  while (wpb > 1 && (d.M + wpb - 1) / wpb < 2 * (int64_t)n_sm) wpb--;
  const size_t smem = per_member * wpb;
  cudaError_t e = cudaFuncSetAttribute(integrate_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((d.M + wpb - 1) / wpb);
  integrate_wide<<<grid, 32 * wpb, smem, st>>>(d, wpb);
  return cudaGetLastError();
}
}  // namespace lwfb
