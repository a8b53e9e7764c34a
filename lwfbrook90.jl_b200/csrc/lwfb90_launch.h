// lwfb90_launch.h — launchers of the kernels that live in lwfb90_member.cu (a separate translation unit,
// see the note at the top of that file).
#pragma once
#include <cuda_runtime.h>
#include "lwfb90_device.cuh"

namespace lwfb {
void launch_member_integrate(const DevData& d, cudaStream_t st);
void launch_member_rhs(const DevData& d, double t, const double* swati, const double* gwat, const double* frozen,
                       const double* ntfli_prev, double* du);
}  // namespace lwfb
