// nxc_fold_policy.cuh -- reduction policies (sum / prod / max / min) for nxc_fold.cuh.
// Semantics: accumulate in the dtype's compute type, store once; signed sums / products wrap
// in the unsigned width; float max / min propagate NaN; bool max / min are or / and
// (reference: nx_c_fold.c:63-101, 261-314).
#pragma once
#include "nxc_ops.cuh"
#include "nxc_fold.cuh"

// ---- reduction policies ------------------------------------------------------------------
template <class C, int CLS> struct Lim;
template <> struct Lim<float, NXC_CLS_FLOAT> { __device__ static float lo() { return -INFINITY; } __device__ static float hi() { return INFINITY; } };
template <> struct Lim<double, NXC_CLS_FLOAT> { __device__ static double lo() { return -INFINITY; } __device__ static double hi() { return INFINITY; } };
template <> struct Lim<int32_t, NXC_CLS_SINT> { __device__ static int32_t lo() { return INT32_MIN; } __device__ static int32_t hi() { return INT32_MAX; } };
template <> struct Lim<int64_t, NXC_CLS_SINT> { __device__ static int64_t lo() { return INT64_MIN; } __device__ static int64_t hi() { return INT64_MAX; } };
template <> struct Lim<uint32_t, NXC_CLS_UINT> { __device__ static uint32_t lo() { return 0; } __device__ static uint32_t hi() { return UINT32_MAX; } };
template <> struct Lim<uint64_t, NXC_CLS_UINT> { __device__ static uint64_t lo() { return 0; } __device__ static uint64_t hi() { return UINT64_MAX; } };
template <> struct Lim<uint32_t, NXC_CLS_BOOL> { __device__ static uint32_t lo() { return 0; } __device__ static uint32_t hi() { return 1; } };

template <int OP, int DT> struct RedP {
  typedef DT_<DT> D;
  typedef typename D::S S;
  typedef typename D::S SO;
  typedef typename D::C A;
  static constexpr int cls = D::cls;
  static constexpr bool ok = (OP == NXC_SUM || OP == NXC_PROD)
                                 ? (cls != NXC_CLS_BOOL)
                                 : (cls != NXC_CLS_COMPLEX);
  __device__ __forceinline__ static A identity() {
    if constexpr (cls == NXC_CLS_COMPLEX) {
      return zmk<A>(OP == NXC_PROD ? 1 : 0, 0);
    } else if constexpr (OP == NXC_SUM) {
      return (A)0;
    } else if constexpr (OP == NXC_PROD) {
      return (A)1;
    } else if constexpr (OP == NXC_RMAX) {
      return Lim<A, cls>::lo();
    } else {
      return Lim<A, cls>::hi();
    }
  }
  __device__ __forceinline__ static void step(A &acc, S s, int64_t) { acc = combine(acc, D::ld(s)); }
  // f16: a vector's worth of elements is widened in bulk (nxc_ld_many: hardware converts, NaN
  // patterns redone exactly) -- the per-element software NaN test held f16 row sums at half the HBM rate
  // 8-bit integer sums: four elements per DP4A against a vector of ones (the accumulator is the
  // 32-bit compute type; sums wrap modulo 2^32 and are truncated on store, as the reference's wider
  // accumulator is -- nx_c.h:109-116)
  static constexpr bool BYTE_SUM = OP == NXC_SUM && (DT == NXC_I8 || DT == NXC_U8);
  static constexpr bool MANY = (DT == NXC_F16) || BYTE_SUM;
  template <int N>
  __device__ __forceinline__ static void step_many(A &acc, const S (&vals)[N], int64_t, int64_t) {
    if constexpr (BYTE_SUM && N % 4 == 0) {
#pragma unroll
      for (int j = 0; j < N / 4; j++) {
        const uint32_t w = (uint32_t)(uint8_t)vals[4 * j] | ((uint32_t)(uint8_t)vals[4 * j + 1] << 8) |
                           ((uint32_t)(uint8_t)vals[4 * j + 2] << 16) | ((uint32_t)(uint8_t)vals[4 * j + 3] << 24);
        if constexpr (DT == NXC_I8) acc = (A)__dp4a((int)w, (int)0x01010101, (int)acc);
        else acc = (A)__dp4a(w, 0x01010101u, (unsigned)acc);
      }
    } else {
      typename D::C c[N];
      nxc_ld_many<DT, N>(vals, c);
#pragma unroll
      for (int i = 0; i < N; i++) acc = combine(acc, c[i]);
    }
  }
  __device__ __forceinline__ static A combine(A a, A b) {
    if constexpr (cls == NXC_CLS_COMPLEX) {
      return OP == NXC_SUM ? zadd(a, b) : zmul(a, b);
    } else if constexpr (OP == NXC_SUM) {
      if constexpr (cls == NXC_CLS_SINT) return (A)((typename UT<A>::U)a + (typename UT<A>::U)b);
      else return a + b;
    } else if constexpr (OP == NXC_PROD) {
      if constexpr (cls == NXC_CLS_SINT) return (A)((typename UT<A>::U)a * (typename UT<A>::U)b);
      else return a * b;
    } else if constexpr (cls == NXC_CLS_FLOAT) {
      // NaN sticks (reference: nx_c_fold.c:80-89)
      if constexpr (sizeof(A) == 4) {
        // one instruction: max.NaN / min.NaN return NaN if either input is NaN
        float r;
        if (OP == NXC_RMAX) asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
        else asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
        return r;
      } else {
        // no max.NaN for f64: branch-free all the same -- the IEEE max / min (which drops a NaN),
        // the sum (NaN iff either is), and one unordered compare choosing between them; the
        // branchy form (test a, test b, compare, select) held short f64 rows at 0.63 of the roofline
        double r;
        if (OP == NXC_RMAX)
          asm("{ .reg .pred p; .reg .f64 m, s; max.f64 m, %1, %2; add.f64 s, %1, %2; setp.nan.f64 p, %1, %2; "
              "selp.f64 %0, s, m, p; }" : "=d"(r) : "d"(a), "d"(b));
        else
          asm("{ .reg .pred p; .reg .f64 m, s; min.f64 m, %1, %2; add.f64 s, %1, %2; setp.nan.f64 p, %1, %2; "
              "selp.f64 %0, s, m, p; }" : "=d"(r) : "d"(a), "d"(b));
        return r;
      }
    } else {
      return OP == NXC_RMAX ? (b > a ? b : a) : (b < a ? b : a);
    }
  }
  __device__ __forceinline__ static SO finish(A a) { return D::st(a); }
};


template <int OP, int DT> struct NxcFoldFewLanes<RedP<OP, DT>> {
  static constexpr bool v = (OP == NXC_RMAX || OP == NXC_RMIN) && sizeof(typename RedP<OP, DT>::S) == 8 &&
                            RedP<OP, DT>::cls == NXC_CLS_FLOAT;
};

nxc_status nxc_reduce_sumprod(nxc_ctx *ctx, int op, int dt, const NxcFoldPlan &p);
nxc_status nxc_reduce_maxmin(nxc_ctx *ctx, int op, int dt, const NxcFoldPlan &p);

#define NXC_RED_CASE(OPC)                                                          \
  case OPC: {                                                                      \
    NXC_DISPATCH_DTYPE(dt, { st = NxcMaybeFold<RedP<OPC, DT>, RedP<OPC, DT>::ok>::go(ctx, p); }) \
  } break;
