/*
 * lwfb90_elem.h — elementary functions of the liblwfb90 kernels, in portable C.
 *
 * The sm_100a kernels evaluate exp / log / pow / sin / cos / tan / asin / acos / atan with the
 * routines below, which use nothing but IEEE-754 +, -, *, /, sqrt, fma, rint and bit manipulation.
 * The same text compiles for the device (nvcc) and for any C99 host compiler, and gives
 * bit-identical results on both (nvcc -fmad=false; host -ffp-contract=off; explicit fma()).
 *
 * Why this exists: the reference algorithm (embedded Runge-Kutta + saturation clamps + flux limiters)
 * is numerically chaotic for some soils — a 1-ulp change in a single pow() changes accept/reject
 * decisions and, weeks later, the 4th digit of the solution (see DESIGN.md "Parity").  A tolerance
 * comparison cannot tell such sensitivity from a bug.  With these functions a host can replay a
 * member with exactly the device's arithmetic (the CPU oracle does so when built with
 * -DLWFB_ORACLE_ELEM), and the comparison becomes bit-exact for every member.
 *
 * Accuracy: <= 2 ulp for exp/log/sin/cos/atan on the ranges the model uses; pow(x,y) =
 * exp(y*log(x)) carries a relative error of about |y*log(x)| * 1.2e-16.
 * Domain shortcuts (all that the model needs): pow requires x >= 0; trigonometric arguments are
 * assumed |x| < 1e5.
 */
#ifndef LWFB90_ELEM_H
#define LWFB90_ELEM_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LWFB_ELEM __host__ __device__ __forceinline__
#define LWFB_ELEM_COLD static __host__ __device__ __noinline__
#else
#define LWFB_ELEM static inline
#define LWFB_ELEM_COLD static __attribute__((noinline, unused))
#endif

LWFB_ELEM int64_t lwfb_d2i(double x) {
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(x);
#else
  int64_t i;
  memcpy(&i, &x, 8);
  return i;
#endif
}
LWFB_ELEM double lwfb_i2d(int64_t i) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(i);
#else
  double x;
  memcpy(&x, &i, 8);
  return x;
#endif
}

#define LWFB_LN2_HI 0x1.62e42fee00000p-1
#define LWFB_LN2_LO 0x1.a39ef35793c76p-33
#define LWFB_INVLN2 0x1.71547652b82fep+0
#define LWFB_PIO2_HI 0x1.921fb54442d18p+0
#define LWFB_PIO2_LO 0x1.1a62633145c07p-54
#define LWFB_PIO2_LO2 -0x1.f1976b7ed8fbcp-110
#define LWFB_2OPI 0x1.45f306dc9c883p-1

/* Polynomial coefficients.  On the device they live in constant memory so that they are read as
 * constant-bank operands of the DFMA instructions (a literal would cost two UMOV instructions each);
 * on the host they are ordinary static tables.  Same values, same order of operations. */
#define LWFB_EXP_COEF  { 1.0 / 87178291200.0, 1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, \
                         1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, \
                         1.0 / 24.0, 1.0 / 6.0, 0.5, LWFB_INVLN2, LWFB_LN2_HI, LWFB_LN2_LO, \
                         709.782712893384, -745.1332191019412 }
#define LWFB_LOG_COEF  { 2.0 / 25.0, 2.0 / 23.0, 2.0 / 21.0, 2.0 / 19.0, 2.0 / 17.0, 2.0 / 15.0, 2.0 / 13.0, 2.0 / 11.0, \
                         2.0 / 9.0, 2.0 / 7.0, 2.0 / 5.0, 2.0 / 3.0, LWFB_LN2_HI, LWFB_LN2_LO, 0x1.6a09e667f3bcdp+0 }
#if defined(__CUDACC__)
static __constant__ double lwfb_exp_coef_d[18] = LWFB_EXP_COEF;
static __constant__ double lwfb_log_coef_d[15] = LWFB_LOG_COEF;
#endif
static const double lwfb_exp_coef_h[18] = LWFB_EXP_COEF;
static const double lwfb_log_coef_h[15] = LWFB_LOG_COEF;
#if defined(__CUDA_ARCH__)
#define LWFB_EC(i) lwfb_exp_coef_d[i]
#define LWFB_LGC(i) lwfb_log_coef_d[i]
#else
#define LWFB_EC(i) lwfb_exp_coef_h[i]
#define LWFB_LGC(i) lwfb_log_coef_h[i]
#endif

/* e^x: k = rint(x/ln2), Cody-Waite reduction with fma, degree-14 Taylor polynomial on |r| <= ln2/2, result scaled by
 * 2^k in two exact steps (so that overflow to inf and gradual underflow to 0 come out of the multiplications).
 * Straight-line code (no branches): independent calls interleave in the FP64 pipe.  |x| >= 747 (incl. inf, NaN) is
 * decided on the high word with integer instructions and answered by a select: NaN -> NaN, +big -> inf, -big -> 0. */
LWFB_ELEM double lwfb_exp(double x) {
  const int big = (int)((uint64_t)lwfb_d2i(x) >> 32 & 0x7fffffffu) >= 0x40875800;   /* |x| >= 747 or inf / NaN */
  double kf = rint(x * LWFB_EC(13));
  double r = fma(-kf, LWFB_EC(14), x);
  r = fma(-kf, LWFB_EC(15), r);
  double q = LWFB_EC(0);                    /* 1/14! */
  q = fma(q, r, LWFB_EC(1));                /* 1/13! */
  q = fma(q, r, LWFB_EC(2));
  q = fma(q, r, LWFB_EC(3));
  q = fma(q, r, LWFB_EC(4));
  q = fma(q, r, LWFB_EC(5));
  q = fma(q, r, LWFB_EC(6));
  q = fma(q, r, LWFB_EC(7));
  q = fma(q, r, LWFB_EC(8));
  q = fma(q, r, LWFB_EC(9));
  q = fma(q, r, LWFB_EC(10));
  q = fma(q, r, LWFB_EC(11));
  q = fma(q, r, LWFB_EC(12));
  double res = 1.0 + fma(r * r, q, r);
  int k = big ? 0 : (int)kf;                /* |k| <= 1078 */
  int adj = k > 1023 ? 1023 : (k < -1021 ? -1000 : 0);   /* adj = 0 multiplies by 1.0 (exact) */
  res = res * lwfb_i2d((int64_t)(adj + 1023) << 52);
  k -= adj;
  res = res * lwfb_i2d((int64_t)(k + 1023) << 52);
  const double special = !(x <= 0.0) ? x * INFINITY : 0.0;   /* NaN -> NaN, x > 0 -> inf, else 0 */
  return big ? special : res;
}

/* log(x) for positive, normal, finite x: x = 2^e * m, m in (sqrt(1/2), sqrt(2)]; log(m) = 2 atanh(s),
 * s = (m-1)/(m+1), evaluated as f - (f^2/2 - s*(f^2/2 + R(s^2))) to avoid the rounding error of the division near
 * m = 1.  Straight-line code. */
LWFB_ELEM double lwfb_log_core(double x, int e0) {
  int64_t ix = lwfb_d2i(x);
  int e = e0 + (int)((ix >> 52) & 0x7ff) - 1023;
  ix = (ix & 0x000fffffffffffffLL) | 0x3ff0000000000000LL;
  double m = lwfb_i2d(ix);
  int big = m > LWFB_LGC(14);
  m = big ? m * 0.5 : m;
  e += big;
  double f = m - 1.0;
  double s = f / (2.0 + f);
  double z = s * s;
  double R = LWFB_LGC(0);
  R = fma(R, z, LWFB_LGC(1));
  R = fma(R, z, LWFB_LGC(2));
  R = fma(R, z, LWFB_LGC(3));
  R = fma(R, z, LWFB_LGC(4));
  R = fma(R, z, LWFB_LGC(5));
  R = fma(R, z, LWFB_LGC(6));
  R = fma(R, z, LWFB_LGC(7));
  R = fma(R, z, LWFB_LGC(8));
  R = fma(R, z, LWFB_LGC(9));
  R = fma(R, z, LWFB_LGC(10));
  R = fma(R, z, LWFB_LGC(11));
  R = R * z;
  double hfsq = 0.5 * f * f;
  double dk = (double)e;
  return dk * LWFB_LGC(12) - ((hfsq - (s * (hfsq + R) + dk * LWFB_LGC(13))) - f);
}
LWFB_ELEM double lwfb_log_main(double x) { return lwfb_log_core(x, 0); }
/* everything lwfb_log_main does not cover: 0, subnormal, negative, inf, NaN (rare; kept out of line on the device) */
LWFB_ELEM int lwfb_log_isspecial(double x) {
  return ((uint32_t)((uint64_t)lwfb_d2i(x) >> 32) - 0x00100000u) >= 0x7fe00000u;   /* not in [2^-1022, inf) */
}
LWFB_ELEM_COLD double lwfb_log_cold(double x) {
  if (x != x) return x;
  if (x < 0.0) return NAN;
  if (x == 0.0) return -INFINITY;
  if (x == INFINITY) return x;
  return lwfb_log_core(x * 0x1p54, -54);    /* subnormal */
}
LWFB_ELEM double lwfb_log(double x) {
  double r = lwfb_log_main(x);
  if (lwfb_log_isspecial(x)) r = lwfb_log_cold(x);
  return r;
}

/* x^y for x >= 0 (x < 0 gives NaN; the reference would throw a DomainError there).
 * exp(y*log(x)) already yields NaN for NaN arguments and for x < 0; only y == 0 needs a select. */
LWFB_ELEM double lwfb_pow(double x, double y) {
  double r = lwfb_exp(y * lwfb_log(x));
  return y == 0.0 ? 1.0 : r;
}
/* x^y when lx = lwfb_log(x) is already known (bit-identical to lwfb_pow) */
LWFB_ELEM double lwfb_pow_from_log(double x, double lx, double y) {
  (void)x;
  double r = lwfb_exp(y * lx);
  return y == 0.0 ? 1.0 : r;
}
/* the same for exponents known to be non-zero (saves the select) */
LWFB_ELEM double lwfb_pow_from_log_nz(double lx, double y) { return lwfb_exp(y * lx); }


/* ---- interleaved evaluations -------------------------------------------------------------------------------------
 * lwfb_exp2 / lwfb_exp3 / lwfb_log_main2 evaluate two or three INDEPENDENT arguments with the operations of the
 * chains written alternately, so that the dependent FMA chains of the polynomials overlap in the FP64 pipe (the
 * compiler keeps each inlined call as one serial chain otherwise).  Every chain performs exactly the operations of
 * the scalar function: the results are bit-identical to lwfb_exp / lwfb_log_main. */
#if defined(__CUDACC__)
#define LWFB_UNROLL _Pragma("unroll")
#else
#define LWFB_UNROLL
#endif
#define LWFB_EXP_HEAD(x, big, kf, r, q) \
  const int big = (int)((uint64_t)lwfb_d2i(x) >> 32 & 0x7fffffffu) >= 0x40875800; \
  double kf = rint(x * LWFB_EC(13)); \
  double r = fma(-kf, LWFB_EC(14), x); \
  r = fma(-kf, LWFB_EC(15), r); \
  double q = LWFB_EC(0);
#define LWFB_EXP_TAIL(x, big, kf, r, q, out) { \
  double res = 1.0 + fma(r * r, q, r); \
  int k = big ? 0 : (int)kf; \
  int adj = k > 1023 ? 1023 : (k < -1021 ? -1000 : 0); \
  res = res * lwfb_i2d((int64_t)(adj + 1023) << 52); \
  k -= adj; \
  res = res * lwfb_i2d((int64_t)(k + 1023) << 52); \
  const double special = !(x <= 0.0) ? x * INFINITY : 0.0; \
  out = big ? special : res; }
LWFB_ELEM void lwfb_exp2(double x1, double x2, double* o1, double* o2) {
  LWFB_EXP_HEAD(x1, b1, k1, r1, q1)
  LWFB_EXP_HEAD(x2, b2, k2, r2, q2)
  LWFB_UNROLL
  for (int j = 1; j <= 12; j++) { q1 = fma(q1, r1, LWFB_EC(j)); q2 = fma(q2, r2, LWFB_EC(j)); }
  LWFB_EXP_TAIL(x1, b1, k1, r1, q1, *o1)
  LWFB_EXP_TAIL(x2, b2, k2, r2, q2, *o2)
}
LWFB_ELEM void lwfb_exp3(double x1, double x2, double x3, double* o1, double* o2, double* o3) {
  LWFB_EXP_HEAD(x1, b1, k1, r1, q1)
  LWFB_EXP_HEAD(x2, b2, k2, r2, q2)
  LWFB_EXP_HEAD(x3, b3, k3, r3, q3)
  LWFB_UNROLL
  for (int j = 1; j <= 12; j++) { q1 = fma(q1, r1, LWFB_EC(j)); q2 = fma(q2, r2, LWFB_EC(j)); q3 = fma(q3, r3, LWFB_EC(j)); }
  LWFB_EXP_TAIL(x1, b1, k1, r1, q1, *o1)
  LWFB_EXP_TAIL(x2, b2, k2, r2, q2, *o2)
  LWFB_EXP_TAIL(x3, b3, k3, r3, q3, *o3)
}
#define LWFB_LOG_HEAD(x, e, f, s, z, R) \
  int e; double f, s, z, R; { \
  int64_t ix = lwfb_d2i(x); \
  e = (int)((ix >> 52) & 0x7ff) - 1023; \
  ix = (ix & 0x000fffffffffffffLL) | 0x3ff0000000000000LL; \
  double m = lwfb_i2d(ix); \
  int big = m > LWFB_LGC(14); \
  m = big ? m * 0.5 : m; \
  e += big; \
  f = m - 1.0; \
  s = f / (2.0 + f); \
  z = s * s; \
  R = LWFB_LGC(0); }
#define LWFB_LOG_TAIL(e, f, s, z, R, out) { \
  double Rz = R * z; \
  double hfsq = 0.5 * f * f; \
  double dk = (double)e; \
  out = dk * LWFB_LGC(12) - ((hfsq - (s * (hfsq + Rz) + dk * LWFB_LGC(13))) - f); }
LWFB_ELEM void lwfb_log_main2(double x1, double x2, double* o1, double* o2) {
  LWFB_LOG_HEAD(x1, e1, f1, s1, z1, R1)
  LWFB_LOG_HEAD(x2, e2, f2, s2, z2, R2)
  LWFB_UNROLL
  for (int j = 1; j <= 11; j++) { R1 = fma(R1, z1, LWFB_LGC(j)); R2 = fma(R2, z2, LWFB_LGC(j)); }
  LWFB_LOG_TAIL(e1, f1, s1, z1, R1, *o1)
  LWFB_LOG_TAIL(e2, f2, s2, z2, R2, *o2)
}

/* sin and cos kernels on |r| <= pi/4 (Taylor, exact coefficients 1/n!) */
LWFB_ELEM double lwfb_ksin(double r) {
  double z = r * r;
  double p = -1.0 / 121645100408832000.0;            /* -1/19! */
  p = fma(p, z, 1.0 / 355687428096000.0);            /*  1/17! */
  p = fma(p, z, -1.0 / 1307674368000.0);
  p = fma(p, z, 1.0 / 6227020800.0);
  p = fma(p, z, -1.0 / 39916800.0);
  p = fma(p, z, 1.0 / 362880.0);
  p = fma(p, z, -1.0 / 5040.0);
  p = fma(p, z, 1.0 / 120.0);
  p = fma(p, z, -1.0 / 6.0);
  return fma(r * z, p, r);
}
LWFB_ELEM double lwfb_kcos(double r) {
  double z = r * r;
  double p = -1.0 / 6402373705728000.0;              /* -1/18! */
  p = fma(p, z, 1.0 / 20922789888000.0);             /*  1/16! */
  p = fma(p, z, -1.0 / 87178291200.0);
  p = fma(p, z, 1.0 / 479001600.0);
  p = fma(p, z, -1.0 / 3628800.0);
  p = fma(p, z, 1.0 / 40320.0);
  p = fma(p, z, -1.0 / 720.0);
  p = fma(p, z, 1.0 / 24.0);
  p = fma(p, z, -0.5);
  return fma(z, p, 1.0);
}
LWFB_ELEM double lwfb_reduce_pio2(double x, int* q) {
  double kf = rint(x * LWFB_2OPI);
  double r = fma(-kf, LWFB_PIO2_HI, x);
  r = fma(-kf, LWFB_PIO2_LO, r);
  r = fma(-kf, LWFB_PIO2_LO2, r);
  *q = ((int)kf) & 3;
  return r;
}
LWFB_ELEM double lwfb_sin(double x) {
  int q;
  double r = lwfb_reduce_pio2(x, &q);
  double s = lwfb_ksin(r), c = lwfb_kcos(r);
  return q == 0 ? s : (q == 1 ? c : (q == 2 ? -s : -c));
}
LWFB_ELEM double lwfb_cos(double x) {
  int q;
  double r = lwfb_reduce_pio2(x, &q);
  double s = lwfb_ksin(r), c = lwfb_kcos(r);
  return q == 0 ? c : (q == 1 ? -s : (q == 2 ? -c : s));
}
LWFB_ELEM double lwfb_tan(double x) { return lwfb_sin(x) / lwfb_cos(x); }

/* atan: |x| > 1 -> pi/2 - atan(1/|x|); then subtract the nearest table point c = j/8:
 * atan(a) = atan(c) + atan((a-c)/(1+a*c)), |(a-c)/(1+a*c)| <= 1/16, 9-term alternating series */
LWFB_ELEM double lwfb_atan(double x) {
  const double ATAN_HI[9] = {0x0.0p+0, 0x1.fd5ba9aac2f6ep-4, 0x1.f5b75f92c80ddp-3, 0x1.6f61941e4def1p-2,
                             0x1.dac670561bb4fp-2, 0x1.1e00babdefeb4p-1, 0x1.4978fa3269ee1p-1,
                             0x1.700a7c5784634p-1, 0x1.921fb54442d18p-1};
  const double ATAN_LO[9] = {0x0.0p+0, -0x1.cd37686760c17p-59, 0x1.8ab6e3cf7afbdp-57, -0x1.c63aae6f6e918p-56,
                             0x1.a2b7f222f65e2p-56, -0x1.928df287a668fp-58, 0x1.2419a87f2a458p-56,
                             -0x1.8c34d25aadef6p-56, 0x1.1a62633145c07p-55};
  if (x != x) return x;
  double ax = fabs(x);
  int inv = 0;
  if (ax > 1.0) { ax = 1.0 / ax; inv = 1; }
  int j = (int)rint(ax * 8.0);
  double c = (double)j * 0.125;
  double t = (ax - c) / fma(ax, c, 1.0);
  double z = t * t;
  double p = -1.0 / 19.0;
  p = fma(p, z, 1.0 / 17.0);
  p = fma(p, z, -1.0 / 15.0);
  p = fma(p, z, 1.0 / 13.0);
  p = fma(p, z, -1.0 / 11.0);
  p = fma(p, z, 1.0 / 9.0);
  p = fma(p, z, -1.0 / 7.0);
  p = fma(p, z, 1.0 / 5.0);
  p = fma(p, z, -1.0 / 3.0);
  double at = fma(t * z, p, t);
  double res = ATAN_HI[j] + (at + ATAN_LO[j]);
  if (inv) res = LWFB_PIO2_HI - (res - LWFB_PIO2_LO);
  return x < 0.0 ? -res : res;
}
LWFB_ELEM double lwfb_asin(double x) { return lwfb_atan(x / sqrt((1.0 - x) * (1.0 + x))); }
LWFB_ELEM double lwfb_acos(double x) { return 2.0 * lwfb_atan(sqrt((1.0 - x) / (1.0 + x))); }

#endif /* LWFB90_ELEM_H */
