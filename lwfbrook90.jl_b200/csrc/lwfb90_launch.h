// lwfb90_launch.h — launchers of the kernels that live in their own translation units (lwfb90_member.cu, see the note
// at the top of that file; lwfb90_wide.cu).
#pragma once
#include <cuda_runtime.h>
#include "lwfb90_device.cuh"

namespace lwfb {
void launch_member_integrate(const DevData& d, cudaStream_t st);
void launch_member_rhs(const DevData& d, double t, const double* swati, const double* gwat, const double* frozen,
                       const double* ntfli_prev, double* du);
// lwfb90_wide.cu: warp-per-member kernel for N > 32 (any layer count whose member block fits the shared memory of an SM)
int wide_max_layers(int device);
cudaError_t launch_wide_integrate(const DevData& d, int device, cudaStream_t st);
}  // namespace lwfb
