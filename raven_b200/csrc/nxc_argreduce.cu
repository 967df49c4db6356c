// nxc_argreduce.cu -- argmax / argmin along one axis, int32 result.
// Replaces caml_nx_c_argmax / caml_nx_c_argmin (reference: nx_c_fold.c:832-833,
// nx_c_engine.c:1214-1254, 1454-1473). Semantics (nx_c_fold.c:93-101, 180-198):
// strict comparison so ties keep the FIRST index; the first NaN wins. The
// (value, index) pair makes the combine associative and order-free, so the
// split-and-fold kernels of nxc_fold.cuh give the sequential answer exactly.
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
  // An accumulator that has taken nothing yet has idx < 0. Floats start from NaN so that the
  // one-instruction "greater or unordered" test sends the first element (and any NaN) to the
  // exact rule below; integers start from the weakest value, never take an equal element, and
  // an output that ends with idx < 0 saw only that value: its answer is index 0 (finish).
  __device__ __forceinline__ static A identity() {
    A a;
    a.idx = -1;
    if constexpr (cls == NXC_CLS_FLOAT) a.v = (C)NAN;
    else if constexpr (cls == NXC_CLS_BOOL) a.v = IS_MAX ? (C)0 : (C)1;
    else a.v = IS_MAX ? std::numeric_limits<C>::lowest() : std::numeric_limits<C>::max();
    return a;
  }
  // one element, visited in increasing index order per accumulator: a strict
  // comparison keeps the first of equals, the NaN clause lets the first NaN win
  __device__ __forceinline__ static void step(A &acc, S s, int64_t r) {
    const C v = D::ld(s);
    if constexpr (cls == NXC_CLS_FLOAT) {
      if (!(IS_MAX ? (v <= acc.v) : (v >= acc.v))) {  // v better, or v / acc NaN (acc NaN: empty or stuck)
        const bool take = acc.idx < 0 || (IS_MAX ? (v > acc.v) : (v < acc.v)) || ((v != v) && !(acc.v != acc.v));
        if (take) { acc.v = v; acc.idx = (int32_t)r; }
      }
    } else {
      if (IS_MAX ? (v > acc.v) : (v < acc.v)) { acc.v = v; acc.idx = (int32_t)r; }
    }
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

extern "C" nxc_status nxc_argreduce(nxc_ctx *ctx, int is_max, const nxc_tensor *out,
                                    const nxc_tensor *in, int axis) {
  nxc_status s;
  if ((s = nxc_check_tensor(in)) || (s = nxc_check_tensor(out))) goto fail;
  {
    const int dt = in->dtype;
    const int cls = nxc_dtype_class(dt);
    NxcFoldPlan p;
    if ((s = nxc_fold_plan(in, out, &axis, 1, nxc_elem_size(dt), 4, &p))) goto fail;
    if (cls & NXC_CLS_PACKED) { s = NXC_ERR_PACKED; goto fail; }
    if ((cls & NXC_CLS_COMPLEX) || out->dtype != NXC_I32) { s = NXC_ERR_UNSUPPORTED_DTYPE; goto fail; }
    const int64_t axis_len = in->shape[axis];
    if (axis_len == 0) { s = NXC_ERR_EMPTY_REDUCE; goto fail; }
    if (axis_len > INT32_MAX) { s = NXC_ERR_ARGREDUCE_CAP; goto fail; }
    if (p.O == 0) return NXC_OK;
    nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
    if (is_max) {
      NXC_DISPATCH_DTYPE(dt, { st = NxcMaybeFold<ArgP<1, DT>, ArgP<1, DT>::ok>::go(ctx, p); })
    } else {
      NXC_DISPATCH_DTYPE(dt, { st = NxcMaybeFold<ArgP<0, DT>, ArgP<0, DT>::ok>::go(ctx, p); })
    }
    s = st;
    if (s) goto fail;
    return NXC_OK;
  }
fail:
  if (s && strcmp(s, NXC_ERR_CUDA) != 0) snprintf(ctx->err, sizeof ctx->err, "%s", s);
  return s;
}
