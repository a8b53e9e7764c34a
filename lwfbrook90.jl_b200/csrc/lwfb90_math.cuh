// lwfb90_math.cuh — transcendental call sites of the hot loop.
//
// All pow/exp/log of the RHS and of the isotope step go through these wrappers.  They map onto the
// portable elementary functions of include/lwfb90_elem.h (bit-identical on host and device), which
// are specialised for what the model needs: one log(a) per layer is shared by the three powers of
// the wetness a (a^(-1/m), a^(n/(n-1)), a^l), and log(K) is evaluated once per layer for both
// interfaces that use it.
#pragma once
#include "../../include/lwfb90_elem.h"

namespace lwfb {
__device__ __forceinline__ double m_log(double x) { return lwfb_log(x); }
__device__ __forceinline__ double m_exp(double x) { return lwfb_exp(x); }
__device__ __forceinline__ double m_pow(double x, double y) { return lwfb_pow(x, y); }
// a^e where la = m_log(a) is already known
__device__ __forceinline__ double m_exp_of_scaled_log(double e, double la, double a) {
  return lwfb_pow_from_log(a, la, e);
}
}  // namespace lwfb
