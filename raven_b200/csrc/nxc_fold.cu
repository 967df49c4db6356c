// nxc_fold.cu -- reduce_sum / reduce_prod / reduce_max / reduce_min.
// Replaces caml_nx_c_reduce_* (reference: nx_c_fold.c:824-831) with the funnel
// and driver checks of nx_c_engine.c:1392-1452, 1052-1105.
#include "nxc_fold_policy.cuh"
#include "nxc_map_groups.cuh"

// ---- the plan (shared with argreduce) -------------------------------------------------
nxc_status nxc_fold_plan(const nxc_tensor *in, const nxc_tensor *out, const int *axes, int n_axes,
                         int64_t in_esize, int64_t out_esize, NxcFoldPlan *p) {
  // order-independent bounds/dup check, then squeeze (nx_c_engine.c:1392-1420)
  if (n_axes < 0 || n_axes > in->ndim) return NXC_ERR_AXES;
  bool red[NXC_MAX_NDIM];
  for (int a = 0; a < in->ndim; a++) red[a] = false;
  for (int i = 0; i < n_axes; i++) {
    int a = axes[i];
    if (a < 0 || a >= in->ndim || red[a]) return NXC_ERR_AXES;
    red[a] = true;
  }
  const int kept = in->ndim - n_axes;
  int64_t ostride[NXC_MAX_NDIM];
  if (out->ndim == kept) {
    for (int j = 0; j < kept; j++) ostride[j] = out->strides[j];
  } else if (out->ndim == in->ndim) {
    int j = 0;
    for (int a = 0; a < in->ndim; a++)
      if (!red[a]) { ostride[j] = out->strides[a]; j++; }
  } else {
    return NXC_ERR_OUT_RANK;
  }
  // strictly increasing (nx_c_engine.c:1085-1089)
  for (int i = 1; i < n_axes; i++)
    if (axes[i] <= axes[i - 1]) return NXC_ERR_AXES;

  p->in_base = (const char *)in->data + in->offset * in_esize;
  p->out_base = (char *)out->data + out->offset * out_esize;
  p->O = 1;
  p->R = 1;
  // kept dims: drop size-1, merge where in AND out strides compose
  int nk = 0, j = 0;
  for (int a = 0; a < in->ndim; a++) {
    if (red[a]) continue;
    const int64_t s = in->shape[a], si = in->strides[a], so = ostride[j];
    j++;
    p->O *= s;
    if (s == 1) continue;
    if (nk > 0 && p->k_in[nk - 1] == si * s && p->k_out[nk - 1] == so * s) {
      p->kshape[nk - 1] *= s;
      p->k_in[nk - 1] = si;
      p->k_out[nk - 1] = so;
    } else {
      p->kshape[nk] = s; p->k_in[nk] = si; p->k_out[nk] = so; nk++;
    }
  }
  p->nk = nk;
  // reduced dims: drop size-1, sort by |stride| descending, merge where strides compose
  int nr = 0;
  int64_t rs[NXC_MAX_NDIM], ri[NXC_MAX_NDIM];
  for (int a = 0; a < in->ndim; a++) {
    if (!red[a]) continue;
    p->R *= in->shape[a];
    if (in->shape[a] == 1) continue;
    rs[nr] = in->shape[a]; ri[nr] = in->strides[a]; nr++;
  }
  for (int x = 1; x < nr; x++) {  // insertion sort, stable
    int64_t s = rs[x], t = ri[x];
    int64_t at = t < 0 ? -t : t;
    int y = x - 1;
    while (y >= 0 && (ri[y] < 0 ? -ri[y] : ri[y]) < at) { rs[y + 1] = rs[y]; ri[y + 1] = ri[y]; y--; }
    rs[y + 1] = s; ri[y + 1] = t;
  }
  int m = 0;
  for (int x = 0; x < nr; x++) {
    if (m > 0 && p->r_in[m - 1] == ri[x] * rs[x]) {
      p->rshape[m - 1] *= rs[x];
      p->r_in[m - 1] = ri[x];
    } else {
      p->rshape[m] = rs[x]; p->r_in[m] = ri[x]; m++;
    }
  }
  p->nr = m;
  return NXC_OK;
}

static void identity_bytes(int op, int dt, void *buf) {
  memset(buf, 0, 16);
  if (op != NXC_PROD) return;
  switch (dt) {
    case NXC_F16: { uint16_t v = 0x3C00; memcpy(buf, &v, 2); } break;
    case NXC_BF16: { uint16_t v = 0x3F80; memcpy(buf, &v, 2); } break;
    case NXC_F8E4M3: { uint8_t v = 0x38; memcpy(buf, &v, 1); } break;
    case NXC_F8E5M2: { uint8_t v = 0x3C; memcpy(buf, &v, 1); } break;
    case NXC_F32: case NXC_C32: { float v = 1.0f; memcpy(buf, &v, 4); } break;
    case NXC_F64: case NXC_C64: { double v = 1.0; memcpy(buf, &v, 8); } break;
    default: { uint8_t v = 1; memcpy(buf, &v, 1); } break;  // little-endian integer 1
  }
}

nxc_status nxc_reduce_sumprod(nxc_ctx *ctx, int op, int dt, const NxcFoldPlan &p) {
  nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
  switch (op) { NXC_RED_CASE(NXC_SUM) NXC_RED_CASE(NXC_PROD) default: break; }
  return st;
}

extern "C" nxc_status nxc_reduce(nxc_ctx *ctx, int op, const nxc_tensor *out, const nxc_tensor *in,
                                 const int *axes, int n_axes) {
  NXC_TRACE(ctx, "nxc_reduce");
  nxc_status s;
  if ((s = nxc_check_tensor(in)) || (s = nxc_check_tensor(out))) goto fail;
  if (op < 0 || op >= NXC_REDUCE_COUNT) { s = NXC_ERR_BAD_OP; goto fail; }
  {
    const int dt = in->dtype;
    const int cls = nxc_dtype_class(dt);
    if (cls & NXC_CLS_PACKED) { s = NXC_ERR_PACKED; goto fail; }
    const bool arith = (op == NXC_SUM || op == NXC_PROD);
    if ((arith && (cls & NXC_CLS_BOOL)) || (!arith && (cls & NXC_CLS_COMPLEX)) || out->dtype != dt) {
      s = NXC_ERR_UNSUPPORTED_DTYPE;
      goto fail;
    }
    if (n_axes > NXC_MAX_NDIM) { s = NXC_ERR_NDIM; goto fail; }
    NxcFoldPlan p;
    const int64_t es = nxc_elem_size(dt);
    if ((s = nxc_fold_plan(in, out, axes, n_axes, es, es, &p))) goto fail;
    if (p.O == 0) return NXC_OK;
    if (p.R == 0) {
      if (!arith) { s = NXC_ERR_EMPTY_REDUCE; goto fail; }
      // empty reduced extent: store the identity (reference: nx_c.h:489-495)
      nxc_tensor o2 = *out;
      if (out->ndim != in->ndim - n_axes) {  // keepdims form: same elements, fill as-is
      }
      char ident[16];
      identity_bytes(op, dt, ident);
      const nxc_tensor *ops[1] = {&o2};
      NxcMapPlan mp;
      if ((s = nxc_map_plan(ops, 1, &es, &mp))) goto fail;
      s = nxc_fill_group(ctx, (int)es, mp, ident);
      if (s) goto fail;
      return NXC_OK;
    }
    s = arith ? nxc_reduce_sumprod(ctx, op, dt, p) : nxc_reduce_maxmin(ctx, op, dt, p);
    if (s) goto fail;
    return NXC_OK;
  }
fail:
  if (s && strcmp(s, NXC_ERR_CUDA) != 0) snprintf(ctx->err, sizeof ctx->err, "%s", s);
  return s;
}
