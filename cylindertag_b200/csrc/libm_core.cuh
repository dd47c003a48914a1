// Float elementary functions with the rounding behaviour of the C library the reference links against.
//
// Why: cv::fitLine (called by the reference at corner_detector.cpp:136-163,358) evaluates cosf/sinf/expf, and its
// DIST_WELSCH variant then picks the best of 20 random restarts with `err < min_err`.  A one-ulp difference in any of
// those calls can flip that discrete choice and move a quad corner by up to ~0.05 px, far above the 1e-3 px parity
// bound.  CUDA's cosf/sinf/expf round differently from glibc's, so the kernels carry their own versions that follow
// the published algorithm of glibc >= 2.28 (sinf/cosf/expf by Szabolcs Nagy, originally ARM Optimized Routines):
// double-precision polynomial on a reduced argument, one final rounding to float.  tests/test_fit_core.py checks them
// bit for bit against the host libm over millions of arguments.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CT_HD __host__ __device__ __forceinline__
#else
#define CT_HD inline
#endif

namespace ctag {
namespace core {

CT_HD uint64_t as_u64(double d) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t u;
  memcpy(&u, &d, 8);
  return u;
#endif
}
CT_HD double as_f64(uint64_t u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double d;
  memcpy(&d, &u, 8);
  return d;
#endif
}
CT_HD uint32_t as_u32(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}

// sin/cos polynomial on [-pi/4, pi/4] evaluated in double (coefficients of the sincosf table).  `neg_cos` selects the
// second table entry, whose cosine coefficients are negated (quadrants 2 and 3).
CT_HD float sincos_poly(double x, double x2, int n, bool neg_cos) {
  const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
               c4 = 0x1.99343027bf8c3p-16;
  const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
  if ((n & 1) == 0) {
    double x3 = x * x2;
    double t1 = s2 + x2 * s3;
    double x7 = x3 * x2;
    double s = x + x3 * s1;
    return (float)(s + x7 * t1);
  } else {
    double x4 = x2 * x2;
    double t2 = c3 + x2 * c4;
    double t1 = c0 + x2 * c1;
    double x6 = x4 * x2;
    double c = t1 + x4 * c2;
    double v = c + x6 * t2;
    return (float)(neg_cos ? -v : v);
  }
}

// quadrant reduction for pi/4 <= |x| < 120
CT_HD double sincos_reduce(double x, int* np) {
  const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
  double r = x * hpi_inv;
  int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return x - n * hpi;
}

CT_HD uint32_t abstop12(float x) { return (as_u32(x) >> 20) & 0x7ff; }

// valid for |y| < 120 (the callers pass angles in [-pi, pi])
CT_HD float libm_sinf(float y) {
  double x = y;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
    if (abstop12(y) < abstop12(0x1p-12f)) return y;
    return sincos_poly(x, x * x, 0, false);
  }
  int n;
  x = sincos_reduce(x, &n);
  double s = ((n + 1) & 2) ? -1.0 : 1.0;  // sign table {1,-1,-1,1}[n & 3]
  return sincos_poly(x * s, x * x, n, (n & 2) != 0);
}

CT_HD float libm_cosf(float y) {
  double x = y;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
    if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
    return sincos_poly(x, x * x, 1, false);
  }
  int n;
  x = sincos_reduce(x, &n);
  double s = ((n + 1) & 2) ? -1.0 : 1.0;
  return sincos_poly(x * s, x * x, n ^ 1, (n & 2) != 0);
}

// 2^(i/32) as IEEE doubles minus (i << 47): the exp2f table (values are the correctly rounded 2^(i/32)).
// Kept in global memory on the device: a function-local array indexed per lane would be rebuilt on the stack on
// every call (measured: ~45 % of the quad kernel's instructions were those stores).
#define CTAG_EXP2F_TAB \
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, \
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, \
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull, \
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, \
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull, \
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, \
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, \
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull
#if defined(__CUDACC__)
static __device__ const uint64_t kExp2fTabDev[32] = {CTAG_EXP2F_TAB};
#endif
static const uint64_t kExp2fTabHost[32] = {CTAG_EXP2F_TAB};
CT_HD uint64_t exp2f_tab(int i) {
#ifdef __CUDA_ARCH__
  return kExp2fTabDev[i];
#else
  return kExp2fTabHost[i];
#endif
}

CT_HD float libm_expf(float x) {
  const double N = 32.0;
  const double InvLn2N = 0x1.71547652b82fep+0 * N;
  const double SHIFT = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5 / N / N / N, C1 = 0x1.ebfce50fac4f3p-3 / N / N, C2 = 0x1.62e42ff0c52d6p-1 / N;
  double xd = (double)x;
  uint32_t abstop = (as_u32(x) >> 20) & 0x7ff;
  if (abstop >= ((as_u32(88.0f) >> 20) & 0x7ff)) {
    if (as_u32(x) == as_u32(-INFINITY)) return 0.0f;
    if (abstop >= ((as_u32(INFINITY) >> 20) & 0x7ff)) return x + x;
    if (x > 0x1.62e42ep6f) return INFINITY;
    if (x < -0x1.9fe368p6f) return 0.0f;
  }
  double z = InvLn2N * xd;
  double kd = z + SHIFT;
  uint64_t ki = as_u64(kd);
  kd -= SHIFT;
  double r = z - kd;
  uint64_t t = exp2f_tab((int)(ki % 32));
  t += ki << (52 - 5);
  double s = as_f64(t);
  z = C0 * r + C1;
  double r2 = r * r;
  double y = C2 * r + 1;
  y = z * r2 + y;
  y = y * s;
  return (float)y;
}

// atan2f: the reference evaluates `atan2(float,float) * 180 / CV_PI` (float atan2f, float product, double division).
// glibc 2.39's atan2f is the old fdlibm routine (not correctly rounded, and CPU-variant dependent); its result only
// feeds angle comparisons against 5/10/50 degree thresholds, a 6-element sort and the final edge normal, where a
// one-ulp difference moves nothing by more than 1e-5 px.  The device uses CUDA's atan2f.
CT_HD float atan2_f(float y, float x) { return atan2f(y, x); }
CT_HD double atan2_deg(float y, float x) { return (double)(atan2_f(y, x) * 180.0f) / 3.1415926535897932384626433832795; }

}  // namespace core
}  // namespace ctag
