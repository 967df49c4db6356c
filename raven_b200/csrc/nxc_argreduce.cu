// nxc_argreduce.cu -- argmax / argmin along one axis, int32 result.
// Replaces caml_nx_c_argmax / caml_nx_c_argmin (reference: nx_c_fold.c:832-833,
// nx_c_engine.c:1214-1254, 1454-1473). Semantics (nx_c_fold.c:93-101, 180-198):
// strict comparison so ties keep the FIRST index; the first NaN wins. The
// (value, index) pair makes the combine associative and order-free, so the
// split-and-fold kernels of nxc_fold.cuh give the sequential answer exactly.
#include "nxc_arg_policy.cuh"

extern "C" nxc_status nxc_argreduce(nxc_ctx *ctx, int is_max, const nxc_tensor *out,
                                    const nxc_tensor *in, int axis) {
  NXC_TRACE(ctx, "nxc_argreduce");
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
