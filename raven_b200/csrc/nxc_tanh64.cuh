// nxc_tanh64.cuh -- float64 tanh within 2 ulp of the reference's.
//
// The reference's f64 tanh is the host libm's (nx_c_map.c:1199 -> glibc tanh): the classical
// Sun formulation -- tanh through expm1 of +-2|x|, expm1 by reduction x = k ln2 + r with a
// two-part ln2, a degree-5 correction polynomial in r*r/2 evaluated in glibc's split
// (Estrin-style) form, and a k-dependent reconstruction. libdevice's tanh / expm1 land up to
// 3 ulp away from it (round 1, measured on B200). What is written here is that published
// algorithm with the multiply-adds glibc's x86-64 build fuses (its FMA ifunc variant, the one
// every FMA-capable host selects) spelled as explicit fma() and every other operation left a
// separate IEEE operation (the translation units including this header build with
// -fmad=false): 40 M random arguments over [-25, 25] and 2^-60..1 agree with glibc 2.39 to the
// last bit, for tanh and for expm1. On a host without FMA glibc's plain variant differs from
// this one by at most 3 ulp on 0.002 % of arguments.
//
// __host__ __device__: the CPU suite compiles this header with gcc and compares it with the
// host libm (tests/test_oracle_tanh64.py); the GPU suite compares the kernel with the oracle.
#pragma once

#include <stdint.h>
#include <string.h>
#include <math.h>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#endif

namespace nxc_t64 {

__host__ __device__ inline uint64_t bits(double x) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
__host__ __device__ inline double from_bits(uint64_t u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double x; memcpy(&x, &u, 8); return x;
#endif
}
// y * 2^k for results that stay normal
__host__ __device__ inline double scale2(double y, int k) { return from_bits(bits(y) + ((uint64_t)(int64_t)k << 52)); }

// e^x - 1 for finite |x| < 56 ln2 (tanh never asks for more)
__host__ __device__ inline double expm1_core(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
               invln2 = 1.44269504088896338700e+00;
  const double Q1 = -3.33333333333331316428e-02, Q2 = 1.58730158725481460165e-03, Q3 = -7.93650757867487942473e-05,
               Q4 = 4.00821782732936239552e-06, Q5 = -2.01099218183624371326e-07;
  const uint32_t hx = (uint32_t)(bits(x) >> 32) & 0x7FFFFFFFu;
  const bool negx = (bits(x) >> 63) != 0;
  double hi, lo, c = 0.0;
  int k = 0;
  if (hx > 0x3FD62E42u) {            // |x| > ln2 / 2
    if (hx < 0x3FF0A2B2u) {          // |x| < 1.5 ln2
      if (!negx) { hi = x - ln2_hi; lo = ln2_lo; k = 1; }
      else { hi = x + ln2_hi; lo = -ln2_lo; k = -1; }
    } else {
      k = (int)fma(invln2, x, negx ? -0.5 : 0.5);
      const double t = (double)k;
      hi = fma(-t, ln2_hi, x);
      lo = t * ln2_lo;
    }
    x = hi - lo;
    c = (hi - x) - lo;
  } else if (hx < 0x3C900000u) {     // |x| < 2^-54
    return x;
  }
  const double hfx = 0.5 * x, hxs = x * hfx;
  const double R1 = fma(hxs, Q1, 1.0), h2 = hxs * hxs, R2 = fma(hxs, Q3, Q2), h4 = h2 * h2, R3 = fma(hxs, Q5, Q4);
  const double r1 = fma(h4, R3, fma(h2, R2, R1));
  double t = fma(-r1, hfx, 3.0);
  double e = hxs * ((r1 - t) / fma(-x, t, 6.0));
  if (k == 0) return x - fma(x, e, -hxs);
  e = fma(x, e - c, -c);
  e -= hxs;
  if (k == -1) return 0.5 * (x - e) - 0.5;
  if (k == 1) return x < -0.25 ? -2.0 * (e - (x + 0.5)) : 1.0 + 2.0 * (x - e);
  if (k <= -2 || k > 56) return scale2(1.0 - (e - x), k) - 1.0;
  if (k < 20) {
    t = from_bits((uint64_t)(0x3FF00000u - (0x200000u >> k)) << 32);  // 1 - 2^-k
    return scale2(t - (e - x), k);
  }
  t = from_bits((uint64_t)(uint32_t)((0x3FF - k) << 20) << 32);        // 2^-k
  double y = x - (e + t);
  y += 1.0;
  return scale2(y, k);
}

__host__ __device__ inline double tanh64(double x) {
  const uint64_t b = bits(x);
  const uint32_t ix = (uint32_t)(b >> 32) & 0x7FFFFFFFu;
  const bool neg = (b >> 63) != 0;
  if (ix >= 0x7FF00000u) {  // +-inf -> +-1, NaN -> NaN
    if ((b & 0x000FFFFFFFFFFFFFull) != 0) return x + x;
    return neg ? -1.0 : 1.0;
  }
  double z;
  if (ix < 0x40360000u) {                              // |x| < 22
    if ((b << 1) == 0) return x;                       // +-0
    if (ix < 0x3C800000u) return x * (1.0 + x);        // |x| < 2^-55
    const double ax = from_bits(b & 0x7FFFFFFFFFFFFFFFull);
    if (ix >= 0x3FF00000u) {                           // |x| >= 1
      const double t = expm1_core(2.0 * ax);
      z = 1.0 - 2.0 / (t + 2.0);
    } else {
      const double t = expm1_core(-2.0 * ax);
      z = -t / (t + 2.0);
    }
  } else {
    z = 1.0 - 1.0e-300;                                // rounds to 1
  }
  return neg ? -z : z;
}

}  // namespace nxc_t64
