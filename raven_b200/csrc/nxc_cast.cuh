// nxc_cast.cuh -- the src x dst cast matrix (17 x 17 compute dtypes).
//
// Conversion policy restated from the reference (nx_c_map.c:182-218): a real
// dst takes the real part of a complex src; an int dst SATURATES from a
// float/complex src (NaN -> 0, clamp to range; nx_c.h:234-257) and WRAPS from
// an int/bool src; a complex dst takes a real src as re+0i; a bool dst is
// (v != 0), so NaN -> true. Narrow float dsts round from their float compute
// type, so f64 -> f16 double-rounds through float exactly as the reference.
#pragma once
#include "nxc_map_groups.cuh"

template <class S> struct IntLim;
#define NXC_LIM(T, BITS_, SIGNED_) template <> struct IntLim<T> { static constexpr int bits = BITS_; static constexpr bool sgn = SIGNED_; };
NXC_LIM(int8_t, 8, true) NXC_LIM(uint8_t, 8, false) NXC_LIM(int16_t, 16, true) NXC_LIM(uint16_t, 16, false)
NXC_LIM(int32_t, 32, true) NXC_LIM(uint32_t, 32, false) NXC_LIM(int64_t, 64, true) NXC_LIM(uint64_t, 64, false)
#undef NXC_LIM

// Saturating double -> integer storage type (reference: nx_c.h:234-250).
template <class S> __device__ __forceinline__ S nxc_f2i(double v) {
  constexpr int w = IntLim<S>::bits;
  if (v != v) return (S)0;
  if (IntLim<S>::sgn) {
    const double lim = (double)(1ull << (w - 1));  // 2^(w-1), exact for every width
    if (v <= -lim) return (S)((uint64_t)1 << (w - 1));
    if (v >= lim) return (S)(((uint64_t)1 << (w - 1)) - 1);
    return (S)(int64_t)v;
  } else {
    const double lim = (w == 64) ? 18446744073709551616.0 : (double)(1ull << (w & 63));  // 2^w
    if (v <= 0.0) return (S)0;
    if (v >= lim) return (S)(~(uint64_t)0);
    return (S)(uint64_t)v;
  }
}

template <int SRC, int DST> struct KCast {
  typedef DT_<SRC> A;
  typedef DT_<DST> B;
  static constexpr int NIN = 1;
  typedef typename B::S S0; typedef typename A::S S1; typedef typename A::S S2; typedef typename A::S S3;
  typedef NxcNoP P;
  typedef typename A::C CA;
  typedef typename B::C CB;

  // real part of the loaded source value, in the source compute type
  template <class T> __device__ __forceinline__ static T real_of(T v) { return v; }
  __device__ __forceinline__ static float real_of(cf32 v) { return v.re; }
  __device__ __forceinline__ static double real_of(cf64 v) { return v.re; }

  // float-class destinations: the conversion on compute-type values (see KUn::op)
  static constexpr int OUT_DT = (B::cls == NXC_CLS_FLOAT) ? DST : -1, IN_DT = SRC;
  __device__ __forceinline__ static CB op(CA a, CA, const P &) {
    if constexpr (B::cls == NXC_CLS_FLOAT) return (CB)real_of(a);
    else return CB();
  }
  __device__ __forceinline__ static S0 run(S1 a, S2, S3, const P &) {
    CA v = A::ld(a);
    if constexpr (B::cls == NXC_CLS_FLOAT) {
      return B::st((CB)real_of(v));
    } else if constexpr (B::cls == NXC_CLS_COMPLEX) {
      typedef typename ZR<CB>::R R;
      if constexpr (A::cls == NXC_CLS_COMPLEX) return zmk<CB>((R)v.re, (R)v.im);
      else return zmk<CB>((R)v, (R)0);
    } else if constexpr (B::cls == NXC_CLS_BOOL) {
      if constexpr (A::cls == NXC_CLS_COMPLEX) return bool_s{(uint8_t)((v.re != 0 || v.im != 0) ? 1 : 0)};
      else return bool_s{(uint8_t)((v != 0) ? 1 : 0)};
    } else {  // integer destination
      if constexpr (A::cls == NXC_CLS_FLOAT || A::cls == NXC_CLS_COMPLEX)
        return nxc_f2i<typename B::S>((double)real_of(v));
      else
        return (typename B::S)v;  // modular wrap; sign/zero extension follows the source type
    }
  }
};

#define NXC_CAST_DST_SWITCH(SRCT)                                                          \
  {                                                                                        \
    nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;                                             \
    NXC_DISPATCH_DTYPE(dst, { st = nxc_map_launch<KCast<SRCT, DT>>(ctx, p, NxcNoP{}); })   \
    return st;                                                                             \
  }
