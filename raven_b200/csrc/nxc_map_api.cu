// nxc_map_api.cu -- C-ABI entry points of the map family: validate, plan,
// dispatch. This is the device-side equivalent of the reference's map funnel and
// driver (nx_c_engine.c:1362-1384, 831-869) and of the hand-assembled cmp / cast
// stubs (nx_c_map.c:1231-1280): same checks in the same order, same statuses.
#include <string.h>

#include "nxc_map_groups.cuh"

nxc_status nxc_cast_packed(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in);
nxc_status nxc_copy_packed(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in);

static const int CLS_NUM = NXC_CLS_SINT | NXC_CLS_UINT | NXC_CLS_FLOAT | NXC_CLS_COMPLEX;
static const int CLS_FC = NXC_CLS_FLOAT | NXC_CLS_COMPLEX;
static const int CLS_INTF = NXC_CLS_SINT | NXC_CLS_UINT | NXC_CLS_FLOAT;
static const int CLS_INT = NXC_CLS_SINT | NXC_CLS_UINT;

// Which dtype classes each op's dispatch table fills (reference:
// nx_c_map.c:224-294 and the per-op table definitions).
static int map1_mask(int op) {
  switch (op) {
    case NXC_NEG: case NXC_RECIP: case NXC_ABS: case NXC_SIGN: return CLS_NUM;
    case NXC_ERF: return NXC_CLS_FLOAT;
    case NXC_TRUNC: case NXC_CEIL: case NXC_FLOOR: case NXC_ROUND: return CLS_INTF;
    default: return (op >= 0 && op < NXC_MAP1_COUNT) ? CLS_FC : 0;
  }
}
static int map2_mask(int op) {
  switch (op) {
    case NXC_ADD: case NXC_SUB: case NXC_MUL: case NXC_POW: return CLS_NUM;
    case NXC_IDIV: case NXC_MOD: return CLS_INTF;
    case NXC_FDIV: return CLS_FC;
    case NXC_MAX: case NXC_MIN: return CLS_INTF | NXC_CLS_BOOL;
    case NXC_ATAN2: return NXC_CLS_FLOAT;
    case NXC_XOR: case NXC_OR: case NXC_AND: return CLS_INT | NXC_CLS_BOOL;
    case NXC_SHL: case NXC_SHR: return CLS_INT;
    default: return 0;
  }
}

static nxc_status fail(nxc_ctx *ctx, nxc_status s) {
  if (s && strcmp(s, NXC_ERR_CUDA) != 0) snprintf(ctx->err, sizeof ctx->err, "%s", s);
  return s;
}

static nxc_status check_ops(const nxc_tensor *const *ops, int nop) {
  for (int k = 0; k < nop; k++) {
    nxc_status s = nxc_check_tensor(ops[k]);
    if (s) return s;
  }
  for (int k = 0; k < nop; k++)
    if (nxc_is_packed(ops[k]->dtype)) return NXC_ERR_PACKED;
  return NXC_OK;
}

static nxc_status same_shape(const nxc_tensor *const *ops, int nop) {
  for (int k = 1; k < nop; k++) {
    if (ops[k]->ndim != ops[0]->ndim) return NXC_ERR_RANK_MISMATCH;
    for (int i = 0; i < ops[0]->ndim; i++)
      if (ops[k]->shape[i] != ops[0]->shape[i]) return NXC_ERR_SHAPE;
  }
  return NXC_OK;
}

extern "C" nxc_status nxc_map1(nxc_ctx *ctx, int op, const nxc_tensor *out, const nxc_tensor *a) {
  NXC_TRACE(ctx, "nxc_map1");
  const nxc_tensor *ops[2] = {out, a};
  nxc_status s = check_ops(ops, 2);
  if (s) return fail(ctx, s);
  if (op < 0 || op >= NXC_MAP1_COUNT) return fail(ctx, NXC_ERR_BAD_OP);
  const int dt = out->dtype;
  if (!(map1_mask(op) & nxc_dtype_class(dt)) || a->dtype != dt) return fail(ctx, NXC_ERR_UNSUPPORTED_DTYPE);
  if ((s = same_shape(ops, 2))) return fail(ctx, s);
  const int64_t es[2] = {nxc_elem_size(dt), nxc_elem_size(dt)};
  NxcMapPlan p;
  if ((s = nxc_map_plan(ops, 2, es, &p))) return fail(ctx, s);
  switch (op) {
    case NXC_NEG: case NXC_RECIP: case NXC_ABS: case NXC_SIGN:
    case NXC_TRUNC: case NXC_CEIL: case NXC_FLOOR: case NXC_ROUND:
      s = nxc_map1_group_a(ctx, op, dt, p); break;
    case NXC_SQRT: case NXC_EXP: case NXC_LOG: case NXC_SIN: case NXC_COS: case NXC_TAN:
      s = nxc_map1_group_b(ctx, op, dt, p); break;
    default:
      s = nxc_map1_group_c(ctx, op, dt, p); break;
  }
  return fail(ctx, s);
}

extern "C" nxc_status nxc_map2(nxc_ctx *ctx, int op, const nxc_tensor *out, const nxc_tensor *a,
                               const nxc_tensor *b) {
  NXC_TRACE(ctx, "nxc_map2");
  const nxc_tensor *ops[3] = {out, a, b};
  nxc_status s = check_ops(ops, 3);
  if (s) return fail(ctx, s);
  if (op < 0 || op >= NXC_MAP2_COUNT) return fail(ctx, NXC_ERR_BAD_OP);
  const int dt = out->dtype;
  if (!(map2_mask(op) & nxc_dtype_class(dt)) || a->dtype != dt || b->dtype != dt)
    return fail(ctx, NXC_ERR_UNSUPPORTED_DTYPE);
  if ((s = same_shape(ops, 3))) return fail(ctx, s);
  const int64_t e = nxc_elem_size(dt);
  const int64_t es[3] = {e, e, e};
  NxcMapPlan p;
  if ((s = nxc_map_plan(ops, 3, es, &p))) return fail(ctx, s);
  if (op <= NXC_MOD) s = nxc_map2_group_a(ctx, op, dt, p);
  else s = nxc_map2_group_b(ctx, op, dt, p);
  return fail(ctx, s);
}

extern "C" nxc_status nxc_cmp(nxc_ctx *ctx, int op, const nxc_tensor *out, const nxc_tensor *a,
                              const nxc_tensor *b) {
  NXC_TRACE(ctx, "nxc_cmp");
  const nxc_tensor *ops[3] = {out, a, b};
  nxc_status s = check_ops(ops, 3);
  if (s) return fail(ctx, s);
  if (op < 0 || op >= NXC_CMP_COUNT) return fail(ctx, NXC_ERR_BAD_OP);
  const int dt = a->dtype;  // dispatched on the INPUT dtype
  if (out->dtype != NXC_BOOL || b->dtype != dt) return fail(ctx, NXC_ERR_UNSUPPORTED_DTYPE);
  if ((op == NXC_CMPLT || op == NXC_CMPLE) && (nxc_dtype_class(dt) & NXC_CLS_COMPLEX))
    return fail(ctx, NXC_ERR_UNSUPPORTED_DTYPE);
  if ((s = same_shape(ops, 3))) return fail(ctx, s);
  const int64_t e = nxc_elem_size(dt);
  const int64_t es[3] = {1, e, e};
  NxcMapPlan p;
  if ((s = nxc_map_plan(ops, 3, es, &p))) return fail(ctx, s);
  return fail(ctx, nxc_cmp_group(ctx, op, dt, p));
}

extern "C" nxc_status nxc_where(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *cond,
                                const nxc_tensor *a, const nxc_tensor *b) {
  NXC_TRACE(ctx, "nxc_where");
  const nxc_tensor *ops[4] = {out, cond, a, b};
  nxc_status s = check_ops(ops, 4);
  if (s) return fail(ctx, s);
  const int dt = out->dtype;
  if (cond->dtype != NXC_BOOL || a->dtype != dt || b->dtype != dt)
    return fail(ctx, NXC_ERR_UNSUPPORTED_DTYPE);
  if ((s = same_shape(ops, 4))) return fail(ctx, s);
  const int64_t e = nxc_elem_size(dt);
  const int64_t es[4] = {e, 1, e, e};
  NxcMapPlan p;
  if ((s = nxc_map_plan(ops, 4, es, &p))) return fail(ctx, s);
  return fail(ctx, nxc_where_group(ctx, (int)e, p));
}

extern "C" nxc_status nxc_cast(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *a) {
  NXC_TRACE(ctx, "nxc_cast");
  const nxc_tensor *ops[2] = {out, a};
  nxc_status s;
  for (int k = 0; k < 2; k++)
    if ((s = nxc_check_tensor(ops[k]))) return fail(ctx, s);
  if (nxc_is_packed(out->dtype) || nxc_is_packed(a->dtype)) return fail(ctx, nxc_cast_packed(ctx, out, a));
  if ((s = same_shape(ops, 2))) return fail(ctx, s);
  const int64_t es[2] = {nxc_elem_size(out->dtype), nxc_elem_size(a->dtype)};
  NxcMapPlan p;
  if ((s = nxc_map_plan(ops, 2, es, &p))) return fail(ctx, s);
  return fail(ctx, nxc_cast_group(ctx, a->dtype, out->dtype, p));
}

nxc_status nxc_cast_group(nxc_ctx *ctx, int src, int dst, const NxcMapPlan &p) {
  switch (src) {
    case NXC_F16: case NXC_F32: case NXC_F64: case NXC_BF16: return nxc_cast_group0(ctx, src, dst, p);
    case NXC_F8E4M3: case NXC_F8E5M2: case NXC_I8: case NXC_U8: case NXC_I16:
      return nxc_cast_group1(ctx, src, dst, p);
    case NXC_U16: case NXC_I32: case NXC_U32: case NXC_I64: return nxc_cast_group2(ctx, src, dst, p);
    default: return nxc_cast_group3(ctx, src, dst, p);
  }
}

extern "C" nxc_status nxc_copy(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *a) {
  NXC_TRACE(ctx, "nxc_copy");
  const nxc_tensor *ops[2] = {out, a};
  if (nxc_valid_dtype(out->dtype) && nxc_is_packed(out->dtype) && a->dtype == out->dtype && out->ndim <= NXC_MAX_NDIM &&
      a->ndim <= NXC_MAX_NDIM)
    return fail(ctx, nxc_copy_packed(ctx, out, a));
  nxc_status s = check_ops(ops, 2);
  if (s) return fail(ctx, s);
  if (a->dtype != out->dtype) return fail(ctx, NXC_ERR_UNSUPPORTED_DTYPE);
  if ((s = same_shape(ops, 2))) return fail(ctx, s);
  const int64_t e = nxc_elem_size(out->dtype);
  const int64_t es[2] = {e, e};
  NxcMapPlan p;
  if ((s = nxc_map_plan(ops, 2, es, &p))) return fail(ctx, s);
  return fail(ctx, nxc_copy_group(ctx, (int)e, p));
}

extern "C" nxc_status nxc_fill(nxc_ctx *ctx, const nxc_tensor *out, const void *scalar) {
  NXC_TRACE(ctx, "nxc_fill");
  const nxc_tensor *ops[1] = {out};
  nxc_status s = check_ops(ops, 1);
  if (s) return fail(ctx, s);
  const int64_t e = nxc_elem_size(out->dtype);
  NxcMapPlan p;
  if ((s = nxc_map_plan(ops, 1, &e, &p))) return fail(ctx, s);
  return fail(ctx, nxc_fill_group(ctx, (int)e, p, scalar));
}
