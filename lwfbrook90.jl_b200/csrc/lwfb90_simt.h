// lwfb90_simt.h — the few CUDA built-ins the kernels use, with a host stand-in for each.
//
// Under nvcc this header adds nothing: the kernels call the real intrinsics.
//
// With -DLWFB_SIMT_EMUL (TEST BUILDS ONLY, see tests/simt/) the same kernel text compiles with a host
// C++ compiler: qualifiers vanish, bit casts become memcpy, and the warp-level primitives
// (__shfl*_sync, __ballot_sync, __any_sync, __all_sync) are routed to a lock-step warp emulator that
// runs the 32 lanes of a warp as coroutines and checks that every lane reaches the same primitive
// (a divergent shuffle is undefined behaviour on the GPU and a hard error in the emulator).  The
// product library is never built this way; the emulator exists so that the warp-synchronous logic
// of the kernels can be checked bit-for-bit against the CPU oracle on a machine without a GPU.
#pragma once

#if defined(LWFB_SIMT_EMUL) && !defined(__CUDACC__)
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)

namespace simt {
// implemented by the emulator (tests/simt/simt_emul.cpp)
uint64_t exchange(uint64_t v, int src_lane, int op_id);     // value of src_lane (own value if out of range)
unsigned vote(bool p, int op_id);                           // ballot over the warp
int lane_id();
bool block_or(bool p, int op_id);                           // __syncthreads_or over the emulated thread block
int block_count(bool p, int op_id);                         // __syncthreads_count
void named_sync(int n_threads, int op_id);                  // bar.sync 1, n_threads (counted named barrier)
int warps_per_block();                                      // warps the emulator runs as one thread block
}  // namespace simt

template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
static inline double __longlong_as_double(long long i) { double x; memcpy(&x, &i, 8); return x; }
static inline long long __double_as_longlong(double x) { long long i; memcpy(&i, &x, 8); return i; }
static inline uint32_t __float_as_uint(float x) { uint32_t i; memcpy(&i, &x, 4); return i; }
static inline float __uint_as_float(uint32_t i) { float x; memcpy(&x, &i, 4); return x; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }

template <typename T> static inline uint64_t simt_pack(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> static inline T simt_unpack(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

#define LWFB_OP_ID (__LINE__)
template <typename T> static inline T simt_shfl(T v, int src, int op) { return simt_unpack<T>(simt::exchange(simt_pack(v), src, op)); }
#define __shfl_sync(mask, v, src) simt_shfl((v), (src), LWFB_OP_ID)
#define __shfl_up_sync(mask, v, delta) simt_shfl((v), simt::lane_id() - (int)(delta), LWFB_OP_ID)
#define __shfl_down_sync(mask, v, delta) simt_shfl((v), simt::lane_id() + (int)(delta), LWFB_OP_ID)
#define __shfl_xor_sync(mask, v, o) simt_shfl((v), simt::lane_id() ^ (int)(o), LWFB_OP_ID)
#define __ballot_sync(mask, p) simt::vote((p), LWFB_OP_ID)
#define __any_sync(mask, p) (simt::vote((p), LWFB_OP_ID) != 0u)
#define __all_sync(mask, p) (simt::vote((p), LWFB_OP_ID) == 0xffffffffu)
#define __syncwarp(mask) ((void)simt::vote(true, LWFB_OP_ID))
#define __syncthreads_or(p) (simt::block_or((p) != 0, LWFB_OP_ID) ? 1 : 0)
#define __syncthreads() ((void)simt::block_or(false, LWFB_OP_ID))
#define __syncthreads_count(p) simt::block_count((p) != 0, LWFB_OP_ID)
// redux.sync min over the lanes named in `mask` (every lane passes the mask of its own group)
static inline unsigned simt_reduce_min(unsigned mask, unsigned v, int op) {
  unsigned r = 0xffffffffu;
  for (int l = 0; l < 32; l++) {
    const unsigned vl = (unsigned)simt::exchange(v, l, op * 64 + l);
    if ((mask >> l) & 1u) r = vl < r ? vl : r;
  }
  return r;
}
#define __reduce_min_sync(mask, v) simt_reduce_min((mask), (v), LWFB_OP_ID)
static inline unsigned simt_reduce_max(unsigned mask, unsigned v, int op) {
  unsigned r = 0u;
  for (int l = 0; l < 32; l++) {
    const unsigned vl = (unsigned)simt::exchange(v, l, op * 64 + l);
    if ((mask >> l) & 1u) r = vl > r ? vl : r;
  }
  return r;
}
#define __reduce_max_sync(mask, v) simt_reduce_max((mask), (v), LWFB_OP_ID)
static inline int __double2hiint(double x) { long long i; memcpy(&i, &x, 8); return (int)(i >> 32); }
static inline int __double2loint(double x) { long long i; memcpy(&i, &x, 8); return (int)(i & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo) { long long i = ((long long)hi << 32) | (unsigned)lo; double x; memcpy(&x, &i, 8); return x; }
// value barrier: the optimiser must treat x as unknown from here on
#define LWFB_OPAQUE(x) asm volatile("" : "+x"(x))
// asynchronous 16-byte copy global -> shared (cp.async) and the wait for the thread's outstanding copies
static inline void lwfb_cp_async16(void* smem_dst, const void* gmem_src) { memcpy(smem_dst, gmem_src, 16); }
static inline void lwfb_cp_async_wait_all() {}
#else
#include <cuda_runtime.h>
#define LWFB_OPAQUE(x) asm volatile("" : "+d"(x))
// asynchronous 16-byte copy global -> shared (LDGSTS: the data does not pass through registers and the thread does not
// wait for it) and the wait for the thread's outstanding copies
__device__ __forceinline__ void lwfb_cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void lwfb_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif
