// nxc_ops.cuh -- scalar semantics of every elementwise op, per dtype class.
//
// Normative source: the reference's kernel expressions (nx_c_map.c:305-488 for
// unary, :499-740 for binary, :747-803 for comparisons) and its policy notes
// (nx_c.h:345-363). Only the scalar meaning is restated; how elements are
// walked is nxc_map.cuh's business. Each op is `Op<C, CLS>` with a constexpr
// `ok` (false -> "dtype not supported for this operation") and `f` in the
// compute type C. No --use_fast_math, no FMA contraction (TUs including this
// header are built with -fmad=false) so float results follow the reference's
// plain IEEE expression order.
#pragma once

#include "nxc_common.cuh"
#include "nxc_tanh64.cuh"

// ---- libm by compute type -----------------------------------------------------
#define NXC_M1(name, ffn, dfn)                                         \
  __device__ __forceinline__ float m_##name(float x) { return ffn(x); } \
  __device__ __forceinline__ double m_##name(double x) { return dfn(x); }
NXC_M1(sqrt, sqrtf, sqrt) NXC_M1(exp, expf, exp) NXC_M1(log, logf, log) NXC_M1(sin, sinf, sin)
NXC_M1(cos, cosf, cos) NXC_M1(asin, asinf, asin) NXC_M1(acos, acosf, acos)
NXC_M1(atan, atanf, atan) NXC_M1(cosh, coshf, cosh) NXC_M1(asinh, asinhf, asinh)
// CUDA's tanf / sinhf / tanhf are specified to 4 / 3 / 2 ulp and measured 3 ulp away
// from glibc on B200; the 2-ulp parity bound needs the double-precision routine
// rounded once (these ops are compute-class in the reference too, nx_c_map.c:1193-1199).
__device__ __forceinline__ float m_tan(float x) { return (float)tan((double)x); }
__device__ __forceinline__ double m_tan(double x) { return tan(x); }
__device__ __forceinline__ float m_sinh(float x) { return (float)sinh((double)x); }
__device__ __forceinline__ double m_sinh(double x) { return sinh(x); }
__device__ __forceinline__ float m_tanh(float x) { return (float)tanh((double)x); }
// f64 tanh: libdevice's is 3 ulp from glibc at worst; nxc_tanh64.cuh restates glibc's own
// expm1-based algorithm with its fused multiply-adds explicit (bit-identical on 40 M samples).
__device__ __forceinline__ double m_tanh(double x) { return nxc_t64::tanh64(x); }
NXC_M1(trunc, truncf, trunc) NXC_M1(ceil, ceilf, ceil) NXC_M1(floor, floorf, floor)
NXC_M1(round, roundf, round) NXC_M1(erf, erff, erf) NXC_M1(fabs, fabsf, fabs)
#undef NXC_M1
__device__ __forceinline__ float m_fmod(float a, float b) { return fmodf(a, b); }
__device__ __forceinline__ double m_fmod(double a, double b) { return fmod(a, b); }
__device__ __forceinline__ float m_pow(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double m_pow(double a, double b) { return pow(a, b); }
__device__ __forceinline__ float m_atan2(float a, float b) { return atan2f(a, b); }
__device__ __forceinline__ double m_atan2(double a, double b) { return atan2(a, b); }
__device__ __forceinline__ float m_hypot(float a, float b) { return hypotf(a, b); }
__device__ __forceinline__ double m_hypot(double a, double b) { return hypot(a, b); }
__device__ __forceinline__ float m_nan(float) { return __uint_as_float(0x7FC00000u); }
__device__ __forceinline__ double m_nan(double) { return __longlong_as_double(0x7FF8000000000000LL); }
__device__ __forceinline__ float m_copysign(float a, float b) { return copysignf(a, b); }
__device__ __forceinline__ double m_copysign(double a, double b) { return copysign(a, b); }

// ---- complex helpers ------------------------------------------------------------
template <class Z> struct ZR;
template <> struct ZR<cf32> { typedef float R; };
template <> struct ZR<cf64> { typedef double R; };
template <class Z> __device__ __forceinline__ Z zmk(typename ZR<Z>::R re, typename ZR<Z>::R im) { Z z; z.re = re; z.im = im; return z; }
template <class Z> __device__ __forceinline__ Z zadd(Z a, Z b) { return zmk<Z>(a.re + b.re, a.im + b.im); }
template <class Z> __device__ __forceinline__ Z zsub(Z a, Z b) { return zmk<Z>(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ bool m_isinf(float x) { return isinf(x); }
__device__ __forceinline__ bool m_isinf(double x) { return isinf(x); }
// (a+bi)(c+di) with the C99 Annex G.5.1 recovery the reference's `*` on _Complex performs
// (an infinite operand must give an infinite, not NaN, product). The slow branch runs only
// when both parts came out NaN.
template <class Z> __device__ __forceinline__ Z zmul(Z p, Z q) {
  typedef typename ZR<Z>::R R;
  R a = p.re, b = p.im, c = q.re, d = q.im;
  R x = a * c - b * d, y = a * d + b * c;
  if (x != x && y != y) {
    bool recalc = false;
    if (m_isinf(a) || m_isinf(b)) {
      a = m_copysign(m_isinf(a) ? (R)1 : (R)0, a);
      b = m_copysign(m_isinf(b) ? (R)1 : (R)0, b);
      if (c != c) c = m_copysign((R)0, c);
      if (d != d) d = m_copysign((R)0, d);
      recalc = true;
    }
    if (m_isinf(c) || m_isinf(d)) {
      c = m_copysign(m_isinf(c) ? (R)1 : (R)0, c);
      d = m_copysign(m_isinf(d) ? (R)1 : (R)0, d);
      if (a != a) a = m_copysign((R)0, a);
      if (b != b) b = m_copysign((R)0, b);
      recalc = true;
    }
    if (!recalc && (m_isinf(a * c) || m_isinf(b * d) || m_isinf(a * d) || m_isinf(b * c))) {
      if (a != a) a = m_copysign((R)0, a);
      if (b != b) b = m_copysign((R)0, b);
      if (c != c) c = m_copysign((R)0, c);
      if (d != d) d = m_copysign((R)0, d);
      recalc = true;
    }
    if (recalc) {
      const R inf = (R)INFINITY;
      x = inf * (a * c - b * d);
      y = inf * (a * d + b * c);
    }
  }
  return zmk<Z>(x, y);
}
// Smith's division: the scaled form C compilers use for `a / b` on _Complex.
template <class Z> __device__ __forceinline__ Z zdiv(Z a, Z b) {
  typedef typename ZR<Z>::R R;
  if (m_fabs(b.re) >= m_fabs(b.im)) {
    if (b.re == (R)0 && b.im == (R)0) return zmk<Z>(a.re / b.re, a.im / b.re);
    R r = b.im / b.re, d = b.re + b.im * r;
    return zmk<Z>((a.re + a.im * r) / d, (a.im - a.re * r) / d);
  }
  R r = b.re / b.im, d = b.re * r + b.im;
  return zmk<Z>((a.re * r + a.im) / d, (a.im * r - a.re) / d);
}
template <class Z> __device__ __forceinline__ typename ZR<Z>::R zabs(Z a) { return m_hypot(a.re, a.im); }
template <class Z> __device__ __forceinline__ Z zexp(Z a) {
  typedef typename ZR<Z>::R R;
  R e = m_exp(a.re);
  if (a.im == (R)0) return zmk<Z>(e, a.im);
  return zmk<Z>(e * m_cos(a.im), e * m_sin(a.im));
}
template <class Z> __device__ __forceinline__ Z zlog(Z a) {
  return zmk<Z>(m_log(zabs(a)), m_atan2(a.im, a.re));
}
template <class Z> __device__ __forceinline__ Z zsqrt(Z a) {
  typedef typename ZR<Z>::R R;
  if (a.re == (R)0 && a.im == (R)0) return zmk<Z>((R)0, a.im);
  R m = zabs(a);
  R t = m_sqrt((m + m_fabs(a.re)) * (R)0.5);
  if (a.re >= (R)0) return zmk<Z>(t, a.im / ((R)2 * t));
  return zmk<Z>(m_fabs(a.im) / ((R)2 * t), m_copysign(t, a.im));
}
template <class Z> __device__ __forceinline__ Z zsinh(Z a) {
  return zmk<Z>(m_sinh(a.re) * m_cos(a.im), m_cosh(a.re) * m_sin(a.im));
}
template <class Z> __device__ __forceinline__ Z zcosh(Z a) {
  return zmk<Z>(m_cosh(a.re) * m_cos(a.im), m_sinh(a.re) * m_sin(a.im));
}
template <class Z> __device__ __forceinline__ Z zsin(Z a) {
  return zmk<Z>(m_sin(a.re) * m_cosh(a.im), m_cos(a.re) * m_sinh(a.im));
}
template <class Z> __device__ __forceinline__ Z zcos(Z a) {
  return zmk<Z>(m_cos(a.re) * m_cosh(a.im), -(m_sin(a.re) * m_sinh(a.im)));
}
template <class Z> __device__ __forceinline__ Z ztanh(Z a) { return zdiv(zsinh(a), zcosh(a)); }
template <class Z> __device__ __forceinline__ Z ztan(Z a) { return zdiv(zsin(a), zcos(a)); }
// asin / acos after Kahan ("Branch cuts for complex elementary functions"): built
// from sqrt(1-z) and sqrt(1+z) so the sign of a zero imaginary part picks the
// side of the branch cut, as C99's casin / cacos do. atan z = (i/2)(log(1-iz) - log(1+iz)).
template <class Z> __device__ __forceinline__ Z zasin(Z a) {
  typedef typename ZR<Z>::R R;
  Z s1 = zsqrt(zmk<Z>((R)1 - a.re, -a.im));
  Z s2 = zsqrt(zmk<Z>((R)1 + a.re, a.im));
  return zmk<Z>(m_atan2(a.re, s1.re * s2.re - s1.im * s2.im), m_asinh(s1.re * s2.im - s1.im * s2.re));
}
template <class Z> __device__ __forceinline__ Z zacos(Z a) {
  typedef typename ZR<Z>::R R;
  Z s1 = zsqrt(zmk<Z>((R)1 - a.re, -a.im));
  Z s2 = zsqrt(zmk<Z>((R)1 + a.re, a.im));
  return zmk<Z>((R)2 * m_atan2(s1.re, s2.re), m_asinh(s2.re * s1.im - s2.im * s1.re));
}
template <class Z> __device__ __forceinline__ Z zatan(Z a) {
  typedef typename ZR<Z>::R R;
  Z l1 = zlog(zmk<Z>((R)1 + a.im, -a.re));  // log(1 - i z)
  Z l2 = zlog(zmk<Z>((R)1 - a.im, a.re));   // log(1 + i z)
  Z d = zsub(l1, l2);
  return zmk<Z>(-(R)0.5 * d.im, (R)0.5 * d.re);
}
template <class Z> __device__ __forceinline__ Z zpow(Z a, Z b) {
  typedef typename ZR<Z>::R R;
  if (a.re == (R)0 && a.im == (R)0) {
    if (b.re == (R)0 && b.im == (R)0) return zmk<Z>((R)1, (R)0);
    return zmk<Z>((R)0, (R)0);
  }
  return zexp(zmul(b, zlog(a)));
}

// ---- integer helpers ----------------------------------------------------------------
template <class T> struct UT;
template <> struct UT<int32_t> { typedef uint32_t U; };
template <> struct UT<uint32_t> { typedef uint32_t U; };
template <> struct UT<int64_t> { typedef uint64_t U; };
template <> struct UT<uint64_t> { typedef uint64_t U; };

// base^e by squaring, wrapping; a negative exponent is 0 except |base| == 1
// (reference: nx_c_map.c:147-165).
template <class C> __device__ __forceinline__ C ipow_signed(C base, C e) {
  typedef typename UT<C>::U U;
  if (e < 0) return (base == 1) ? (C)1 : (base == -1) ? ((e & 1) ? (C)-1 : (C)1) : (C)0;
  U b = (U)base, r = 1;
  while (e > 0) { if (e & 1) r *= b; e >>= 1; if (e) b *= b; }
  return (C)r;
}
template <class C> __device__ __forceinline__ C ipow_unsigned(C b, C e) {
  C r = 1;
  while (e > 0) { if (e & 1) r *= b; e >>= 1; if (e) b *= b; }
  return r;
}

// ======================================================================================
// Unary ops
// ======================================================================================
template <int OP, class C, int CLS> struct Un { static constexpr bool ok = false; __device__ static C f(C x) { return x; } };

#define NXC_UN(OP, CLS, EXPR)                                           \
  template <class C> struct Un<OP, C, CLS> {                            \
    static constexpr bool ok = true;                                    \
    __device__ __forceinline__ static C f(C x) { return EXPR; }          \
  };
// neg / recip / abs / sign (reference: nx_c_map.c:305-366)
NXC_UN(NXC_NEG, NXC_CLS_SINT, (C)(-(typename UT<C>::U)x))
NXC_UN(NXC_NEG, NXC_CLS_UINT, (C)(-x))
NXC_UN(NXC_NEG, NXC_CLS_FLOAT, -x)
NXC_UN(NXC_NEG, NXC_CLS_COMPLEX, zmk<C>(-x.re, -x.im))
NXC_UN(NXC_RECIP, NXC_CLS_SINT, (x == 0 ? (C)0 : (C)(1 / x)))
NXC_UN(NXC_RECIP, NXC_CLS_UINT, (x == 0 ? (C)0 : (C)(1 / x)))
NXC_UN(NXC_RECIP, NXC_CLS_FLOAT, (C)1 / x)
NXC_UN(NXC_RECIP, NXC_CLS_COMPLEX, zdiv(zmk<C>(1, 0), x))
NXC_UN(NXC_ABS, NXC_CLS_SINT, (x < 0 ? (C)(-(typename UT<C>::U)x) : x))
NXC_UN(NXC_ABS, NXC_CLS_UINT, x)
NXC_UN(NXC_ABS, NXC_CLS_FLOAT, m_fabs(x))
NXC_UN(NXC_ABS, NXC_CLS_COMPLEX, zmk<C>(zabs(x), 0))
NXC_UN(NXC_SIGN, NXC_CLS_SINT, (C)((x > 0) - (x < 0)))
NXC_UN(NXC_SIGN, NXC_CLS_UINT, (C)(x != 0))
NXC_UN(NXC_SIGN, NXC_CLS_FLOAT, ((x != x) ? x : (C)((x > 0) - (x < 0))))
template <class C> struct Un<NXC_SIGN, C, NXC_CLS_COMPLEX> {
  static constexpr bool ok = true;
  __device__ __forceinline__ static C f(C x) {
    typename ZR<C>::R m = zabs(x);
    if (m == 0) return zmk<C>(0, 0);
    return zmk<C>(x.re / m, x.im / m);
  }
};
// transcendentals: float + complex (reference: nx_c_map.c:388-440)
#define NXC_UN_TRANS(OP, name)                 \
  NXC_UN(OP, NXC_CLS_FLOAT, m_##name(x))       \
  NXC_UN(OP, NXC_CLS_COMPLEX, z##name(x))
NXC_UN_TRANS(NXC_SQRT, sqrt) NXC_UN_TRANS(NXC_EXP, exp) NXC_UN_TRANS(NXC_LOG, log)
NXC_UN_TRANS(NXC_SIN, sin) NXC_UN_TRANS(NXC_COS, cos) NXC_UN_TRANS(NXC_TAN, tan)
NXC_UN_TRANS(NXC_ASIN, asin) NXC_UN_TRANS(NXC_ACOS, acos) NXC_UN_TRANS(NXC_ATAN, atan)
NXC_UN_TRANS(NXC_SINH, sinh) NXC_UN_TRANS(NXC_COSH, cosh) NXC_UN_TRANS(NXC_TANH, tanh)
// erf: float only (reference: nx_c_map.c:443-456)
NXC_UN(NXC_ERF, NXC_CLS_FLOAT, m_erf(x))
// rounding: identity on ints, libm on floats (reference: nx_c_map.c:460-488)
#define NXC_UN_ROUND(OP, name)            \
  NXC_UN(OP, NXC_CLS_SINT, x)             \
  NXC_UN(OP, NXC_CLS_UINT, x)             \
  NXC_UN(OP, NXC_CLS_FLOAT, m_##name(x))
NXC_UN_ROUND(NXC_TRUNC, trunc) NXC_UN_ROUND(NXC_CEIL, ceil) NXC_UN_ROUND(NXC_FLOOR, floor)
NXC_UN_ROUND(NXC_ROUND, round)

// ======================================================================================
// Binary ops
// ======================================================================================
template <int OP, class C, int CLS, int BITS> struct Bin { static constexpr bool ok = false; __device__ static C f(C a, C) { return a; } };

#define NXC_BIN(OP, CLS, EXPR)                                               \
  template <class C, int BITS> struct Bin<OP, C, CLS, BITS> {                \
    static constexpr bool ok = true;                                         \
    __device__ __forceinline__ static C f(C a, C b) { return EXPR; }          \
  };
// add / sub / mul (reference: nx_c_map.c:499-531)
#define NXC_BIN_ARITH(OP, SYM, ZF)                                            \
  NXC_BIN(OP, NXC_CLS_SINT, (C)((typename UT<C>::U)a SYM(typename UT<C>::U) b)) \
  NXC_BIN(OP, NXC_CLS_UINT, (C)(a SYM b))                                     \
  NXC_BIN(OP, NXC_CLS_FLOAT, a SYM b)                                         \
  NXC_BIN(OP, NXC_CLS_COMPLEX, ZF(a, b))
NXC_BIN_ARITH(NXC_ADD, +, zadd) NXC_BIN_ARITH(NXC_SUB, -, zsub) NXC_BIN_ARITH(NXC_MUL, *, zmul)
// idiv (reference: nx_c_map.c:535-553)
NXC_BIN(NXC_IDIV, NXC_CLS_SINT, (b == 0 ? (C)0 : b == -1 ? (C)(-(typename UT<C>::U)a) : (C)(a / b)))
NXC_BIN(NXC_IDIV, NXC_CLS_UINT, (b == 0 ? (C)0 : (C)(a / b)))
NXC_BIN(NXC_IDIV, NXC_CLS_FLOAT, m_trunc(a / b))
// fdiv (reference: nx_c_map.c:556-569)
NXC_BIN(NXC_FDIV, NXC_CLS_FLOAT, a / b)
NXC_BIN(NXC_FDIV, NXC_CLS_COMPLEX, zdiv(a, b))
// mod (reference: nx_c_map.c:573-589)
NXC_BIN(NXC_MOD, NXC_CLS_SINT, (b == 0 ? (C)0 : b == -1 ? (C)0 : (C)(a % b)))
NXC_BIN(NXC_MOD, NXC_CLS_UINT, (b == 0 ? (C)0 : (C)(a % b)))
NXC_BIN(NXC_MOD, NXC_CLS_FLOAT, m_fmod(a, b))
// max / min, NaN-propagating on floats (reference: nx_c_map.c:594-625)
#define NXC_BIN_MM(OP, SYM)                                                          \
  NXC_BIN(OP, NXC_CLS_SINT, (a SYM b ? a : b))                                       \
  NXC_BIN(OP, NXC_CLS_UINT, (a SYM b ? a : b))                                       \
  NXC_BIN(OP, NXC_CLS_BOOL, (a SYM b ? a : b))                                       \
  NXC_BIN(OP, NXC_CLS_FLOAT, ((a != a || b != b) ? m_nan(a) : (a SYM b ? a : b)))
NXC_BIN_MM(NXC_MAX, >) NXC_BIN_MM(NXC_MIN, <)
// pow (reference: nx_c_map.c:629-645)
NXC_BIN(NXC_POW, NXC_CLS_SINT, ipow_signed<C>(a, b))
NXC_BIN(NXC_POW, NXC_CLS_UINT, ipow_unsigned<C>(a, b))
NXC_BIN(NXC_POW, NXC_CLS_FLOAT, m_pow(a, b))
NXC_BIN(NXC_POW, NXC_CLS_COMPLEX, zpow(a, b))
// atan2: float only (reference: nx_c_map.c:648-661)
NXC_BIN(NXC_ATAN2, NXC_CLS_FLOAT, m_atan2(a, b))
// bitwise: ints + bool (reference: nx_c_map.c:665-695)
#define NXC_BIN_BW(OP, SYM)                   \
  NXC_BIN(OP, NXC_CLS_SINT, (C)(a SYM b))     \
  NXC_BIN(OP, NXC_CLS_UINT, (C)(a SYM b))     \
  NXC_BIN(OP, NXC_CLS_BOOL, (C)(a SYM b))
NXC_BIN_BW(NXC_XOR, ^) NXC_BIN_BW(NXC_OR, |) NXC_BIN_BW(NXC_AND, &)
// shifts: a count negative or >= the storage width gives 0 (reference: nx_c_map.c:697-740)
NXC_BIN(NXC_SHL, NXC_CLS_SINT, ((b < 0 || b >= (C)BITS) ? (C)0 : (C)((typename UT<C>::U)a << b)))
NXC_BIN(NXC_SHL, NXC_CLS_UINT, ((b >= (C)BITS) ? (C)0 : (C)(a << b)))
NXC_BIN(NXC_SHR, NXC_CLS_SINT, ((b < 0 || b >= (C)BITS) ? (C)0 : (C)(a >> b)))
NXC_BIN(NXC_SHR, NXC_CLS_UINT, ((b >= (C)BITS) ? (C)0 : (C)(a >> b)))

// ======================================================================================
// Comparisons (bool result) (reference: nx_c_map.c:747-803)
// ======================================================================================
template <int OP, class C, int CLS> struct Cmp {
  static constexpr bool ok = (CLS != NXC_CLS_COMPLEX);
  __device__ __forceinline__ static bool f(C a, C b) {
    if (OP == NXC_CMPEQ) return a == b;
    if (OP == NXC_CMPNE) return a != b;
    if (OP == NXC_CMPLT) return a < b;
    return a <= b;
  }
};
template <int OP, class C> struct Cmp<OP, C, NXC_CLS_COMPLEX> {
  static constexpr bool ok = (OP == NXC_CMPEQ || OP == NXC_CMPNE);
  __device__ __forceinline__ static bool f(C a, C b) {
    bool eq = (a.re == b.re) && (a.im == b.im);
    return OP == NXC_CMPEQ ? eq : !eq;
  }
};

// ======================================================================================
// Kernel-op adapters (storage in, storage out) for nxc_map.cuh
// ======================================================================================
struct NxcNoP {};

template <int OP, int DT> struct KUn {
  typedef DT_<DT> D;
  typedef Un<OP, typename D::C, D::cls> O;
  static constexpr int NIN = 1;
  typedef typename D::S S0; typedef typename D::S S1; typedef typename D::S S2; typedef typename D::S S3;
  typedef NxcNoP P;
  // the op on compute-type values: the vector paths convert f16 inputs and f16 / bf16 results in bulk
  // (nxc_ld_many, nxc_pack16) around it
  static constexpr int OUT_DT = DT, IN_DT = DT;
  __device__ __forceinline__ static typename D::C op(typename D::C a, typename D::C, const P &) { return O::f(a); }
  __device__ __forceinline__ static S0 run(S1 a, S2, S3, const P &) { return D::st(O::f(D::ld(a))); }
};
template <int OP, int DT> struct KBin {
  typedef DT_<DT> D;
  // a transposed operand goes through the shared-memory tiled kernel for the hot arithmetic ops
  static constexpr bool TILED = (OP == NXC_ADD || OP == NXC_SUB || OP == NXC_MUL || OP == NXC_FDIV ||
                                 OP == NXC_MAX || OP == NXC_MIN) && sizeof(typename DT_<DT>::S) <= 8;
  typedef Bin<OP, typename D::C, D::cls, 8 * (int)sizeof(typename D::S)> O;
  static constexpr int NIN = 2;
  typedef typename D::S S0; typedef typename D::S S1; typedef typename D::S S2; typedef typename D::S S3;
  typedef NxcNoP P;
  static constexpr int OUT_DT = DT, IN_DT = DT;
  __device__ __forceinline__ static typename D::C op(typename D::C a, typename D::C b, const P &) { return O::f(a, b); }
  __device__ __forceinline__ static S0 run(S1 a, S2 b, S3, const P &) { return D::st(O::f(D::ld(a), D::ld(b))); }
};
template <int OP, int DT> struct KCmp {
  typedef DT_<DT> D;
  typedef Cmp<OP, typename D::C, D::cls> O;
  static constexpr int NIN = 2;
  typedef bool_s S0; typedef typename D::S S1; typedef typename D::S S2; typedef typename D::S S3;
  typedef NxcNoP P;
  __device__ __forceinline__ static S0 run(S1 a, S2 b, S3, const P &) {
    return bool_s{(uint8_t)(O::f(D::ld(a), D::ld(b)) ? 1 : 0)};
  }
};
