// nxc_move.cu -- pad, cat, gather, scatter.
// Replaces caml_nx_c_pad / _cat / _gather / _scatter (reference:
// nx_c_move.c:203-569).
//
//   pad      border slabs are filled, the interior is one strided copy -- every
//            output element is written exactly once, through the map kernels
//            (same decomposition idea as the reference, nx_c_move.c:203-283).
//   cat      one strided copy per member into its slice (nx_c_move.c:285-318).
//   gather   out[c] = data[c with axis -> idx[c]]; int32 indices, Python-wrapped
//            once, then bounds-checked (nx_c_move.c:342-362).
//   scatter  `Add uses hardware atomics (CAS on the containing word for sub-word
//            dtypes): integer results are exact, float results differ from the
//            reference's serial walk by summation order only. `Set must make the
//            LAST write in row-major order win (nx_c_move.c:444-460): a first pass
//            records the winning source position per target with atomicMax, a
//            second pass lets only winners write, so duplicates are deterministic.
//
// An out-of-range index cannot raise from inside a kernel: it is recorded in a
// device flag and reported ("index out of bounds ...", Failure) by the call
// itself after a stream sync when NX_CUDA_SYNC_CHECKS=1 (the default), or by the
// next nxc_sync / nxc_d2h when it is 0 (the flag is a sticky word of the status page).
#include "nxc_map_groups.cuh"
#include "nxc_fold.cuh"

static nxc_status fail(nxc_ctx *ctx, nxc_status s) {
  if (s && strcmp(s, NXC_ERR_CUDA) != 0) snprintf(ctx->err, sizeof ctx->err, "%s", s);
  return s;
}

static nxc_status copy_into(nxc_ctx *ctx, const nxc_tensor *dst, const nxc_tensor *src) {
  const int64_t e = nxc_elem_size(dst->dtype);
  const int64_t es[2] = {e, e};
  const nxc_tensor *ops[2] = {dst, src};
  NxcMapPlan p;
  nxc_status s = nxc_map_plan(ops, 2, es, &p);
  if (s) return s;
  return nxc_copy_group(ctx, (int)e, p);
}
static nxc_status fill_view(nxc_ctx *ctx, const nxc_tensor *dst, const void *scalar) {
  const int64_t e = nxc_elem_size(dst->dtype);
  const nxc_tensor *ops[1] = {dst};
  NxcMapPlan p;
  nxc_status s = nxc_map_plan(ops, 1, &e, &p);
  if (s) return s;
  return nxc_fill_group(ctx, (int)e, p, scalar);
}

extern "C" nxc_status nxc_pad(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in, const void *fill,
                              const int64_t *before) {
  NXC_TRACE(ctx, "nxc_pad");
  nxc_status s;
  if ((s = nxc_check_tensor(out)) || (s = nxc_check_tensor(in))) return fail(ctx, s);
  if (nxc_is_packed(out->dtype)) return fail(ctx, NXC_ERR_PACKED);
  if (out->ndim != in->ndim || out->dtype != in->dtype) return fail(ctx, NXC_ERR_SHAPE);
  nxc_tensor slab = *out;
  for (int d = 0; d < out->ndim; d++) {
    const int64_t b = before[d], after = out->shape[d] - in->shape[d] - b;
    if (b < 0 || after < 0) return fail(ctx, NXC_ERR_SHAPE);
    if (b > 0) {
      slab.shape[d] = b;
      if ((s = fill_view(ctx, &slab, fill))) return fail(ctx, s);
    }
    if (after > 0) {
      slab.shape[d] = after;
      slab.offset += (b + in->shape[d]) * out->strides[d];
      if ((s = fill_view(ctx, &slab, fill))) return fail(ctx, s);
      slab.offset -= (b + in->shape[d]) * out->strides[d];
    }
    slab.shape[d] = in->shape[d];
    slab.offset += b * out->strides[d];
  }
  return fail(ctx, copy_into(ctx, &slab, in));
}

extern "C" nxc_status nxc_cat(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *const *ins, int n, int axis) {
  NXC_TRACE(ctx, "nxc_cat");
  nxc_status s;
  if ((s = nxc_check_tensor(out))) return fail(ctx, s);
  if (nxc_is_packed(out->dtype)) return fail(ctx, NXC_ERR_PACKED);
  if (axis < 0 || axis >= out->ndim) return fail(ctx, NXC_ERR_AXIS);
  int64_t pos = 0;
  for (int m = 0; m < n; m++) {
    if ((s = nxc_check_tensor(ins[m]))) return fail(ctx, s);
    if (ins[m]->ndim != out->ndim || ins[m]->dtype != out->dtype) return fail(ctx, NXC_ERR_SHAPE);
    nxc_tensor slice = *out;
    slice.offset = out->offset + pos * out->strides[axis];
    for (int d = 0; d < out->ndim; d++) slice.shape[d] = ins[m]->shape[d];
    if ((s = copy_into(ctx, &slice, ins[m]))) return fail(ctx, s);
    pos += ins[m]->shape[axis];
  }
  return NXC_OK;
}

// ---- gather / scatter ----------------------------------------------------------------
struct IdxArgs {
  int ndim, axis;
  int small;
  int64_t total, axis_len;
  NxcFastDiv div[NXC_MAX_NDIM];
  int64_t shape[NXC_MAX_NDIM];
  int64_t s_iter0[NXC_MAX_NDIM];  // strides of the tensor iterated with the index space (out / updates)
  int64_t s_idx[NXC_MAX_NDIM];
  int64_t s_tgt[NXC_MAX_NDIM];    // strides of the tensor addressed through the index (data / out)
};

__device__ __forceinline__ bool idx_offsets(const IdxArgs &a, int64_t it, const int32_t *idx, int64_t &o_iter,
                                            int64_t &o_tgt) {
  int64_t o_idx = 0;
  o_iter = 0;
  o_tgt = 0;
  if (a.small) {
    uint32_t r = (uint32_t)it;
    for (int d = a.ndim - 1; d >= 0; d--) {
      uint32_t q = nxc_fastdiv(r, a.div[d]);
      int64_t c = r - q * a.div[d].d;
      r = q;
      o_iter += c * a.s_iter0[d];
      o_idx += c * a.s_idx[d];
      if (d != a.axis) o_tgt += c * a.s_tgt[d];
    }
  } else {
    int64_t r = it;
    for (int d = a.ndim - 1; d >= 0; d--) {
      int64_t q = r / a.shape[d], c = r - q * a.shape[d];
      r = q;
      o_iter += c * a.s_iter0[d];
      o_idx += c * a.s_idx[d];
      if (d != a.axis) o_tgt += c * a.s_tgt[d];
    }
  }
  int64_t ix = idx[o_idx];
  if (ix < 0) ix += a.axis_len;
  if (ix < 0 || ix >= a.axis_len) return false;
  o_tgt += ix * a.s_tgt[a.axis];
  return true;
}

template <class T>
__global__ void __launch_bounds__(256) gather_kernel(T *__restrict__ out, const T *__restrict__ data,
                                                     const int32_t *__restrict__ idx,
                                                     const __grid_constant__ IdxArgs a, int *oob) {
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int64_t it = (int64_t)blockIdx.x * 256 + threadIdx.x; it < a.total; it += step) {
    int64_t oo, od;
    if (!idx_offsets(a, it, idx, oo, od)) { *oob = 1; continue; }
    out[oo] = data[od];
  }
}

template <class T>
__global__ void __launch_bounds__(256) scatter_set_kernel(T *__restrict__ out, const T *__restrict__ upd,
                                                          const int32_t *__restrict__ idx,
                                                          const __grid_constant__ IdxArgs a, int *oob,
                                                          long long *winner, int pass) {
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int64_t it = (int64_t)blockIdx.x * 256 + threadIdx.x; it < a.total; it += step) {
    int64_t ou, oo;
    if (!idx_offsets(a, it, idx, ou, oo)) { *oob = 1; continue; }
    if (winner == nullptr) out[oo] = upd[ou];
    else if (pass == 0) atomicMax(&winner[oo], (long long)it);
    else if (winner[oo] == (long long)it) out[oo] = upd[ou];
  }
}

// atomic add of one element of dtype DT at `p`
template <int DT> __device__ __forceinline__ void atomic_add_elem(typename DT_<DT>::S *p, typename DT_<DT>::S v) {
  typedef DT_<DT> D;
  typedef typename D::S S;
  if constexpr (DT == NXC_F32) atomicAdd(p, v);
  else if constexpr (DT == NXC_F64) atomicAdd(p, v);
  else if constexpr (DT == NXC_I32) atomicAdd(p, v);
  else if constexpr (DT == NXC_U32) atomicAdd(p, v);
  else if constexpr (DT == NXC_I64 || DT == NXC_U64) atomicAdd((unsigned long long *)p, (unsigned long long)v);
  else if constexpr (DT == NXC_C32) { atomicAdd(&p->re, v.re); atomicAdd(&p->im, v.im); }
  else if constexpr (DT == NXC_C64) { atomicAdd(&p->re, v.re); atomicAdd(&p->im, v.im); }
  else {
    // sub-word dtypes: CAS on the aligned 32-bit word that contains the element
    const uintptr_t addr = (uintptr_t)p;
    unsigned int *word = (unsigned int *)(addr & ~(uintptr_t)3);
    const unsigned int shift = (unsigned int)(addr & 3) * 8;
    const unsigned int mask = (sizeof(S) == 1 ? 0xFFu : 0xFFFFu) << shift;
    unsigned int old = *word, assumed;
    do {
      assumed = old;
      S cur;
      const unsigned int bits = (assumed & mask) >> shift;
      if (sizeof(S) == 1) { uint8_t b = (uint8_t)bits; memcpy(&cur, &b, 1); } else { uint16_t b = (uint16_t)bits; memcpy(&cur, &b, 2); }
      S nv;
      if constexpr (D::cls == NXC_CLS_BOOL) nv = D::st((D::ld(cur) + D::ld(v)) != 0 ? 1u : 0u);
      else nv = D::st(D::ld(cur) + D::ld(v));
      unsigned int nb = 0;
      memcpy(&nb, &nv, sizeof(S));
      old = atomicCAS(word, assumed, (assumed & ~mask) | (nb << shift));
    } while (old != assumed);
  }
}

template <int DT>
__global__ void __launch_bounds__(256) scatter_add_kernel(typename DT_<DT>::S *__restrict__ out,
                                                          const typename DT_<DT>::S *__restrict__ upd,
                                                          const int32_t *__restrict__ idx,
                                                          const __grid_constant__ IdxArgs a, int *oob) {
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int64_t it = (int64_t)blockIdx.x * 256 + threadIdx.x; it < a.total; it += step) {
    int64_t ou, oo;
    if (!idx_offsets(a, it, idx, ou, oo)) { *oob = 1; continue; }
    atomic_add_elem<DT>(out + oo, upd[ou]);
  }
}

static void idx_args(IdxArgs &a, const nxc_tensor *space, const nxc_tensor *idx, const nxc_tensor *tgt, int axis) {
  a.ndim = space->ndim;
  a.axis = axis;
  a.total = nxc_numel(idx);
  a.axis_len = tgt->shape[axis];
  a.small = a.total < 0x7FFFFFFFLL;
  for (int d = 0; d < a.ndim; d++) {
    a.shape[d] = idx->shape[d];
    a.div[d] = nxc_fastdiv_make(a.small ? (uint32_t)idx->shape[d] : 1u);
    a.s_iter0[d] = space->strides[d];
    a.s_idx[d] = idx->strides[d];
    a.s_tgt[d] = tgt->strides[d];
  }
}

static int sync_checks() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("NX_CUDA_SYNC_CHECKS"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}
// The range flag is a word of the context's status page (mapped host memory): sticky, written by
// the kernels, read by the host without a copy. A checked call drains the stream and reports it
// itself; an unchecked one (NX_CUDA_SYNC_CHECKS=0, nxc_gather_trusted, any call inside a captured
// step) leaves it for the next nxc_sync / nxc_d2h.
static nxc_status oob_flag(nxc_ctx *ctx, int **flag) {
  *flag = ctx->dstatus + NXC_ST_INDEX;
  return NXC_OK;
}
static nxc_status oob_check(nxc_ctx *ctx, int *) {
  if (!sync_checks() || nxc_is_capturing(ctx)) return NXC_OK;
  NXC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->hstatus[NXC_ST_INDEX]) {
    ctx->hstatus[NXC_ST_INDEX] = 0;
    return NXC_ERR_INDEX_OOB;
  }
  return NXC_OK;
}
static unsigned grid_for(nxc_ctx *ctx, int64_t total) {
  int64_t b = (total + 255) / 256, cap = (int64_t)ctx->sm_count * 32;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

#define BY_SIZE(es, CALL)                                                         \
  switch (es) {                                                                   \
    case 1: { typedef uint8_t T; CALL; } break;                                   \
    case 2: { typedef uint16_t T; CALL; } break;                                  \
    case 4: { typedef uint32_t T; CALL; } break;                                  \
    case 8: { typedef uint2 T; CALL; } break;                                     \
    default: { typedef uint4 T; CALL; } break;                                    \
  }

static nxc_status gather_impl(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *data, const nxc_tensor *idx, int axis,
                              bool checked);
extern "C" nxc_status nxc_gather(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *data,
                                 const nxc_tensor *idx, int axis) {
  NXC_TRACE(ctx, "nxc_gather");
  return gather_impl(ctx, out, data, idx, axis, true);
}
// The same gather for indices the backend produced itself (argmax / argmin / argsort results,
// as the sharded argreduce reads the local extreme back at the winning index): out-of-range
// indices are still never dereferenced, but the range flag is not read back, so the call does
// not drain the stream.
extern "C" nxc_status nxc_gather_trusted(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *data,
                                         const nxc_tensor *idx, int axis) {
  NXC_TRACE(ctx, "nxc_gather_trusted");
  return gather_impl(ctx, out, data, idx, axis, false);
}
static nxc_status gather_impl(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *data, const nxc_tensor *idx, int axis,
                              bool checked) {
  nxc_status s;
  if ((s = nxc_check_tensor(out)) || (s = nxc_check_tensor(data)) || (s = nxc_check_tensor(idx))) return fail(ctx, s);
  if (nxc_is_packed(out->dtype)) return fail(ctx, NXC_ERR_PACKED);
  if (idx->dtype != NXC_I32 || data->dtype != out->dtype) return fail(ctx, NXC_ERR_UNSUPPORTED_DTYPE);
  if (axis < 0 || axis >= data->ndim) return fail(ctx, NXC_ERR_AXIS);
  if (data->ndim != idx->ndim || data->ndim != out->ndim) return fail(ctx, NXC_ERR_SHAPE);
  for (int d = 0; d < out->ndim; d++)
    if (out->shape[d] != idx->shape[d]) return fail(ctx, NXC_ERR_SHAPE);
  IdxArgs a;
  idx_args(a, out, idx, data, axis);
  if (a.total == 0) return NXC_OK;
  int *flag;
  if ((s = oob_flag(ctx, &flag))) return fail(ctx, s);
  const int64_t es = nxc_elem_size(out->dtype);
  char *ob = (char *)out->data + out->offset * es;
  const char *db = (const char *)data->data + data->offset * es;
  const int32_t *ib = (const int32_t *)idx->data + idx->offset;
  BY_SIZE(es, (gather_kernel<T><<<grid_for(ctx, a.total), 256, 0, ctx->stream>>>((T *)ob, (const T *)db, ib, a, flag)))
  NXC_LAUNCH_CHECK(ctx);
  return checked ? fail(ctx, oob_check(ctx, flag)) : NXC_OK;
}

template <int DT, bool OK> struct ScatterAdd {
  static nxc_status go(nxc_ctx *ctx, char *ob, const char *ub, const int32_t *ib, const IdxArgs &a, int *flag) {
    typedef typename DT_<DT>::S S;
    scatter_add_kernel<DT><<<grid_for(ctx, a.total), 256, 0, ctx->stream>>>((S *)ob, (const S *)ub, ib, a, flag);
    NXC_LAUNCH_CHECK(ctx);
    return NXC_OK;
  }
};

extern "C" nxc_status nxc_scatter(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *idx,
                                  const nxc_tensor *upd, int axis, int mode) {
  NXC_TRACE(ctx, "nxc_scatter");
  nxc_status s;
  if ((s = nxc_check_tensor(out)) || (s = nxc_check_tensor(upd)) || (s = nxc_check_tensor(idx))) return fail(ctx, s);
  const int dt = out->dtype;
  if (nxc_is_packed(dt)) return fail(ctx, NXC_ERR_PACKED);
  if (idx->dtype != NXC_I32 || upd->dtype != dt) return fail(ctx, NXC_ERR_UNSUPPORTED_DTYPE);
  if (axis < 0 || axis >= out->ndim) return fail(ctx, NXC_ERR_AXIS);
  if (out->ndim != idx->ndim || out->ndim != upd->ndim) return fail(ctx, NXC_ERR_SHAPE);
  for (int d = 0; d < out->ndim; d++) {
    if (idx->shape[d] != upd->shape[d]) return fail(ctx, NXC_ERR_SHAPE);
    if (d != axis && idx->shape[d] != out->shape[d]) return fail(ctx, NXC_ERR_SHAPE);
  }
  const bool add = (mode & 1) != 0, unique = (mode & 2) != 0;
  IdxArgs a;
  idx_args(a, upd, idx, out, axis);
  if (a.total == 0) return NXC_OK;
  const int64_t es = nxc_elem_size(dt);
  char *ob = (char *)out->data + out->offset * es;
  const char *ub = (const char *)upd->data + upd->offset * es;
  const int32_t *ib = (const int32_t *)idx->data + idx->offset;
  if (add) {
    int *flag;
    if ((s = oob_flag(ctx, &flag))) return fail(ctx, s);
    nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
    NXC_DISPATCH_DTYPE(dt, { st = ScatterAdd<DT, true>::go(ctx, ob, ub, ib, a, flag); })
    if (st) return fail(ctx, st);
    return fail(ctx, oob_check(ctx, flag));
  }
  // `Set. The winner table is indexed by element offset from `ob`, so it must span
  // the addressed extent of `out` (non-negative strides assumed; otherwise serialise
  // through the unique path after making `out` contiguous is the caller's job).
  int64_t span = 1;
  bool neg = false;
  for (int d = 0; d < out->ndim; d++) {
    if (out->strides[d] < 0) neg = true;
    if (out->shape[d] > 0) span += (out->shape[d] - 1) * (out->strides[d] < 0 ? -out->strides[d] : out->strides[d]);
  }
  long long *winner = nullptr;
  int *flag;
  if (!unique && !neg) {
    void *scr;
    if ((s = nxc_scratch(ctx, (size_t)span * 8, &scr))) return fail(ctx, s);
    if ((s = oob_flag(ctx, &flag))) return fail(ctx, s);
    winner = (long long *)scr;
    NXC_CUDA_TRY(ctx, cudaMemsetAsync(scr, 0xFF, (size_t)span * 8, ctx->stream));  // winner = -1
  } else {
    if ((s = oob_flag(ctx, &flag))) return fail(ctx, s);
  }
  const unsigned g = grid_for(ctx, a.total);
  if (winner) {
    BY_SIZE(es, (scatter_set_kernel<T><<<g, 256, 0, ctx->stream>>>((T *)ob, (const T *)ub, ib, a, flag, winner, 0)))
    NXC_LAUNCH_CHECK(ctx);
    BY_SIZE(es, (scatter_set_kernel<T><<<g, 256, 0, ctx->stream>>>((T *)ob, (const T *)ub, ib, a, flag, winner, 1)))
    NXC_LAUNCH_CHECK(ctx);
  } else {
    BY_SIZE(es, (scatter_set_kernel<T><<<g, 256, 0, ctx->stream>>>((T *)ob, (const T *)ub, ib, a, flag, nullptr, 0)))
    NXC_LAUNCH_CHECK(ctx);
  }
  return fail(ctx, oob_check(ctx, flag));
}
