// lwfb90_math.cuh — FP64 transcendental helpers of the hot loop.
//
// The RHS spends most of its FP64 instructions in pow/exp/log (5 live pow + 2 log + 1 exp per
// layer).  All call sites go through these wrappers so that the implementation can be specialised for
// the argument ranges that actually occur (0 < a <= 1, moderate exponents) without touching the model
// code.  LWFB_EXACT_MATH selects the plain CUDA math library (used for the tightest parity runs).
#pragma once
#include <math.h>

namespace lwfb {

#ifdef LWFB_EXACT_MATH
__device__ __forceinline__ double m_log(double x) { return log(x); }
__device__ __forceinline__ double m_exp(double x) { return exp(x); }
__device__ __forceinline__ double m_pow(double x, double y) { return pow(x, y); }
// a^e where la = log(a) is already known
__device__ __forceinline__ double m_exp_of_scaled_log(double e, double la, double a) { (void)la; return pow(a, e); }
#else
__device__ __forceinline__ double m_log(double x) { return log(x); }
__device__ __forceinline__ double m_exp(double x) { return exp(x); }
__device__ __forceinline__ double m_pow(double x, double y) { return pow(x, y); }
__device__ __forceinline__ double m_exp_of_scaled_log(double e, double la, double a) { (void)la; return pow(a, e); }
#endif

}  // namespace lwfb
