// nxc_arg_policy.cuh -- the (value, index) fold policy of argmax / argmin, shared by the axis
// kernels (nxc_argreduce.cu) and the cross-rank finish of a sharded argreduce (nxc_dist_fold.cu).
// Semantics (reference: nx_c_fold.c:93-101, 180-198): strict comparison so ties keep the FIRST
// index; the first NaN wins.
#pragma once
#include <limits>

#include "nxc_ops.cuh"
#include "nxc_fold.cuh"

template <class C> struct ArgAcc { C v; int32_t idx; };

template <int IS_MAX, int DT> struct ArgP {
  typedef DT_<DT> D;
  typedef typename D::S S;
  typedef int32_t SO;
  typedef typename D::C C;
  typedef ArgAcc<C> A;
  static constexpr int cls = D::cls;
  static constexpr bool ok = (cls != NXC_CLS_COMPLEX);
  // An accumulator that has taken nothing yet has idx < 0 and holds the weakest value (-inf /
  // +inf, the type's lowest / highest): a strict comparison never takes an equal element, so an
  // output that ends with idx < 0 saw nothing but that value and its answer is index 0 (finish).
  __device__ __forceinline__ static A identity() {
    A a;
    a.idx = -1;
    if constexpr (cls == NXC_CLS_FLOAT) a.v = IS_MAX ? (C)-INFINITY : (C)INFINITY;
    else if constexpr (cls == NXC_CLS_BOOL) a.v = IS_MAX ? (C)0 : (C)1;
    else a.v = IS_MAX ? std::numeric_limits<C>::lowest() : std::numeric_limits<C>::max();
    return a;
  }
  // "v beats the accumulator": strictly better, or the first NaN (a NaN accumulator is stuck)
  __device__ __forceinline__ static bool beats(C v, C accv) {
    if constexpr (cls == NXC_CLS_FLOAT) return (IS_MAX ? !(v <= accv) : !(v >= accv)) && (accv == accv);
    else return IS_MAX ? (v > accv) : (v < accv);
  }
  // one element, visited in increasing index order per accumulator
  __device__ __forceinline__ static void step(A &acc, S s, int64_t r) {
    const C v = D::ld(s);
    if (beats(v, acc.v)) { acc.v = v; acc.idx = (int32_t)r; }
  }
  // NaN-propagating extreme of two values
  __device__ __forceinline__ static C better(C a, C b) {
    if constexpr (cls == NXC_CLS_FLOAT && sizeof(C) == 4) {
      float r;
      if (IS_MAX) asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
      else asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
      return r;
    } else if constexpr (cls == NXC_CLS_FLOAT) {
      C r = (IS_MAX ? (b > a) : (b < a)) ? b : a;
      return (b != b) ? b : r;
    } else {
      return (IS_MAX ? (b > a) : (b < a)) ? b : a;
    }
  }
  // N elements at indices r0, r0 + rs, ...: the group's extreme first (N-1 instructions), one
  // comparison against the accumulator, then which element it was (the first one equal to the
  // extreme) -- all predicated, no branch: a thread's walk is often only a few groups long, so
  // "the group beats the accumulator" is not rare enough to branch on (measured: the branching
  // version ran the inner-axis kernels at 0.66-0.77 of the HBM rate, the per-element one at
  // 0.84). Only a NaN in the group -- rare -- takes a branch to the exact first-NaN rule.
  static constexpr bool MANY = true;
  template <int N>
  __device__ __forceinline__ static void step_many(A &acc, const S (&vals)[N], int64_t r0, int64_t rs) {
    C c[N];
#pragma unroll
    for (int i = 0; i < N; i++) c[i] = D::ld(vals[i]);
    C m = c[0];
#pragma unroll
    for (int i = 1; i < N; i++) m = better(m, c[i]);
    const int32_t i0 = (int32_t)r0, is = (int32_t)rs;
    if constexpr (cls == NXC_CLS_FLOAT) {
      if (m != m) {
        if (acc.v == acc.v) {  // the first NaN wins, a NaN accumulator is stuck
          int sel = N - 1;
#pragma unroll
          for (int i = N - 2; i >= 0; i--) sel = (c[i] != c[i]) ? i : sel;
          acc.v = m;
          acc.idx = i0 + sel * is;
        }
        return;
      }
    }
    const bool hit = IS_MAX ? (m > acc.v) : (m < acc.v);  // false for a NaN accumulator
    int sel = N - 1;
#pragma unroll
    for (int i = N - 2; i >= 0; i--) sel = (c[i] == m) ? i : sel;
    acc.v = hit ? m : acc.v;
    acc.idx = hit ? i0 + sel * is : acc.idx;
  }
  // The lanes of one thread group (a power of two <= 32, `mask` their lane mask) hold partials of
  // the same output: the group's extreme by a NaN-propagating butterfly on the VALUE alone, then
  // the lowest index among the lanes that hold it (REDUX.MIN) -- ~15 instructions instead of
  // log2(lanes) merges of (value, index) pairs.
  static constexpr bool WARP = true;
  __device__ __forceinline__ static A warp_combine(A t, int lanes, unsigned mask) {
    C m = t.v;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1)
      if (s < lanes) m = better(m, nxc_shfl_xor(m, s));
    bool mine;
    if constexpr (cls == NXC_CLS_FLOAT) mine = (m != m) ? (t.v != t.v) : (t.v == m);
    else mine = t.v == m;
    const int cand = (mine && t.idx >= 0) ? t.idx : INT32_MAX;
    const int best = __reduce_min_sync(mask, cand);
    A r;
    r.v = m;
    r.idx = best == INT32_MAX ? -1 : best;
    return r;
  }
  __device__ __forceinline__ static A combine(A a, A b) {
    if (a.idx < 0) return b;
    if (b.idx < 0) return a;
    const bool a_first = a.idx < b.idx;
    if constexpr (cls == NXC_CLS_FLOAT) {
      const bool an = a.v != a.v, bn = b.v != b.v;
      if (an || bn) {
        if (an && bn) return a_first ? a : b;
        return an ? a : b;
      }
    }
    const bool a_better = IS_MAX ? (a.v > b.v) : (a.v < b.v);
    const bool b_better = IS_MAX ? (b.v > a.v) : (b.v < a.v);
    if (a_better) return a;
    if (b_better) return b;
    return a_first ? a : b;
  }
  __device__ __forceinline__ static SO finish(A a) { return a.idx < 0 ? 0 : a.idx; }
};
template <int IS_MAX, int DT> struct NxcFoldFewLanes<ArgP<IS_MAX, DT>> { static constexpr bool v = true; };
template <int IS_MAX, int DT> struct NxcFoldPipe<ArgP<IS_MAX, DT>> {
  static constexpr bool v = sizeof(typename ArgP<IS_MAX, DT>::S) >= 8;
};

