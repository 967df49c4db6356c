// nxc_dist_fold.cu -- exchange + fold in ONE kernel over peer memory.
//
// A sharded reduction ends with "combine one small partial per rank". Done with library calls
// that is an all-gather (or NCCL all-reduce) plus a local fold, and for argmax / argmin a chain
// of seven launches and two exchanges (gather the local extreme, offset the index, gather both,
// argmax over ranks, gather the winner). Here the kernel that pushes this rank's partial into the
// peers' mailboxes (nxc_dist.cuh) also waits for theirs and folds all of them, in RANK ORDER on
// every rank, with the backend's own fold policies (nxc_fold_policy.cuh, nxc_arg_policy.cuh):
//
//   nxc_allreduce            sum / prod / max / min of `count` elements, every dtype the
//                            single-device reduce takes; float max / min stay NaN-sticky and every
//                            rank ends with bit-identical results
//   nxc_argreduce_exchange   the finish of an argmax / argmin along the sharded leading axis:
//                            each rank contributes (its local extreme, read from its slab at its
//                            local argreduce's index; that index + the slab's global offset), the
//                            winner is picked by the first-index / first-NaN rule of
//                            nx_c_fold.c:93-101 -- the lowest rank holding the extreme is the
//                            lowest global index, so the answer is the single-device one, exactly.
//
// Payload limits are the mailbox's: 256 KiB per rank for the all-reduce, 16384 outputs for the
// arg-reduce finish; larger ones take the gather-based paths of nxc_dist.cu / sharded.py.
#include "nxc_fold_policy.cuh"
#include "nxc_arg_policy.cuh"
#include "nxc_dist.cuh"

// ---- all-reduce -------------------------------------------------------------------------------
// grid = chunks of 16 KiB; CTA c pushes chunk c of this rank's payload to EVERY rank (itself
// included, so the fold reads all partials the same way), flags each, waits for chunk c of every
// rank and folds it. In place: a CTA reads its chunk of `buf` before it overwrites it.
template <class P>
__global__ void __launch_bounds__(256) nxc_p2p_allreduce_kernel(const NxcP2P a, typename P::S *__restrict__ buf, int64_t count) {
  typedef typename P::S S;
  constexpr int64_t EPC = (int64_t)(NXC_P2P_CHUNK_BYTES / sizeof(S));  // elements per chunk
  const int c = blockIdx.x;
  const uint32_t e = nxc_p2p_epoch(a);
  const int64_t lo = (int64_t)c * EPC;
  const int64_t len = count - lo < EPC ? count - lo : EPC;
  for (int k = 0; k < a.world; k++) {
    int s = a.rank + 1 + k;  // peers first, own mailbox last: the NVLink stores get the head start
    if (s >= a.world) s -= a.world;
    nxc_p2p_copy_out(nxc_p2p_slot(a, s, e, a.rank) + lo * (int64_t)sizeof(S), (const char *)(buf + lo), len * (int64_t)sizeof(S));
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  if ((int)threadIdx.x < a.world) nxc_st_release_sys(nxc_p2p_flag(a, threadIdx.x, e, a.rank, c), e);
  __syncthreads();
  if ((int)threadIdx.x < a.world && !nxc_p2p_wait(nxc_p2p_flag(a, a.rank, e, threadIdx.x, c), e, a.status)) bad = 1;
  __syncthreads();
  if (!bad) {
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
      typename P::A acc = P::identity();
      for (int s = 0; s < a.world; s++)
        P::step(acc, nxc_p2p_load(reinterpret_cast<const S *>(nxc_p2p_slot(a, a.rank, e, s)) + lo + i), s);
      buf[lo + i] = P::finish(acc);
    }
  }
  nxc_p2p_finish(a, e);
}

template <class P, bool OK> struct NxcFusedAllreduce {
  static nxc_status go(nxc_ctx *ctx, void *buf, int64_t count) {
    typedef typename P::S S;
    const int64_t epc = (int64_t)(NXC_P2P_CHUNK_BYTES / sizeof(S));
    const int chunks = (int)((count + epc - 1) / epc);
    nxc_p2p_allreduce_kernel<P><<<chunks, 256, 0, ctx->stream>>>(nxc_p2p_args(ctx), (S *)buf, count);
    NXC_LAUNCH_CHECK(ctx);
    return NXC_OK;
  }
};
template <class P> struct NxcFusedAllreduce<P, false> {
  static nxc_status go(nxc_ctx *, void *, int64_t) { return NXC_ERR_UNSUPPORTED_DTYPE; }
};

nxc_status nxc_p2p_allreduce_fused(nxc_ctx *ctx, void *buf, int64_t count, int dtype, int op) {
  nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
#define NXC_AR_CASE(OPC)                                                                                   \
  case OPC: {                                                                                              \
    NXC_DISPATCH_DTYPE(dtype, { st = NxcFusedAllreduce<RedP<OPC, DT>, RedP<OPC, DT>::ok>::go(ctx, buf, count); }) \
  } break;
  switch (op) { NXC_AR_CASE(NXC_SUM) NXC_AR_CASE(NXC_PROD) NXC_AR_CASE(NXC_RMAX) NXC_AR_CASE(NXC_RMIN) default: break; }
#undef NXC_AR_CASE
  if (st && strcmp(st, NXC_ERR_CUDA) != 0) snprintf(ctx->err, sizeof ctx->err, "%s", st);
  return st;
}

// ---- arg-reduce finish ------------------------------------------------------------------------
struct NxcArgXArgs {
  NxcDimList kept;      // the kept dims of the slab: in_stride = the slab's strides, out_stride unused
  int64_t axis_stride;  // slab stride of the reduced (sharded) axis, elements
  int64_t count;        // outputs
  int64_t idx_offset;   // global index of this slab's first position along the axis
  int small;
};

// grid = chunks of NXC_P2P_ARG_PER_CHUNK outputs; a chunk of the slot holds the values, then the
// int32 global indices.
template <class P>
__global__ void __launch_bounds__(256) nxc_p2p_argreduce_kernel(const NxcP2P a, const __grid_constant__ NxcArgXArgs g,
                                                                const typename P::S *__restrict__ x,
                                                                const int32_t *__restrict__ local_idx, int32_t *__restrict__ out) {
  typedef typename P::S S;
  constexpr int OPC = NXC_P2P_ARG_PER_CHUNK;
  static_assert(OPC * (sizeof(S) + 4) <= NXC_P2P_CHUNK_BYTES, "chunk holds values and indices");
  const int c = blockIdx.x;
  const uint32_t e = nxc_p2p_epoch(a);
  const int64_t lo = (int64_t)c * OPC;
  const int len = (int)(g.count - lo < OPC ? g.count - lo : OPC);
  const size_t chunk_off = (size_t)c * NXC_P2P_CHUNK_BYTES;
  // this rank's partial for its outputs: stays in registers, goes to every rank's mailbox
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    const int64_t o = lo + i;
    const int32_t li = local_idx[o];
    int64_t in_off, out_off;
    nxc_kept_offset(g.kept, o, g.small, in_off, out_off);
    const S v = x[in_off + (int64_t)li * g.axis_stride];
    const int32_t gi = (int32_t)((int64_t)li + g.idx_offset);
    for (int k = 0; k < a.world; k++) {
      int s = a.rank + 1 + k;
      if (s >= a.world) s -= a.world;
      char *slot = nxc_p2p_slot(a, s, e, a.rank) + chunk_off;
      reinterpret_cast<S *>(slot)[i] = v;
      reinterpret_cast<int32_t *>(slot + OPC * sizeof(S))[i] = gi;
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  if ((int)threadIdx.x < a.world) nxc_st_release_sys(nxc_p2p_flag(a, threadIdx.x, e, a.rank, c), e);
  __syncthreads();
  if ((int)threadIdx.x < a.world && !nxc_p2p_wait(nxc_p2p_flag(a, a.rank, e, threadIdx.x, c), e, a.status)) bad = 1;
  __syncthreads();
  if (!bad) {
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      typename P::A acc = P::identity();
      for (int s = 0; s < a.world; s++) {
        const char *slot = nxc_p2p_slot(a, a.rank, e, s) + chunk_off;
        typename P::A p;
        p.v = P::D::ld(nxc_p2p_load(reinterpret_cast<const S *>(slot) + i));
        p.idx = nxc_p2p_load(reinterpret_cast<const int32_t *>(slot + OPC * sizeof(S)) + i);
        acc = P::combine(acc, p);
      }
      out[lo + i] = P::finish(acc);
    }
  }
  nxc_p2p_finish(a, e);
}

template <class P, bool OK> struct NxcFusedArg {
  static nxc_status go(nxc_ctx *ctx, const NxcArgXArgs &g, const void *x, const int32_t *li, int32_t *out) {
    const int chunks = (int)((g.count + NXC_P2P_ARG_PER_CHUNK - 1) / NXC_P2P_ARG_PER_CHUNK);
    nxc_p2p_argreduce_kernel<P><<<chunks, 256, 0, ctx->stream>>>(nxc_p2p_args(ctx), g, (const typename P::S *)x, li, out);
    NXC_LAUNCH_CHECK(ctx);
    return NXC_OK;
  }
};
template <class P> struct NxcFusedArg<P, false> {
  static nxc_status go(nxc_ctx *, const NxcArgXArgs &, const void *, const int32_t *, int32_t *) { return NXC_ERR_UNSUPPORTED_DTYPE; }
};

extern "C" int64_t nxc_argreduce_exchange_max_outputs(nxc_ctx *ctx) {
  return ctx->p2p ? (int64_t)NXC_P2P_MAX_CHUNKS * NXC_P2P_ARG_PER_CHUNK : 0;
}

extern "C" nxc_status nxc_argreduce_exchange(nxc_ctx *ctx, int is_max, const nxc_tensor *out, const nxc_tensor *x,
                                             const nxc_tensor *local_idx, int axis, int64_t idx_offset) {
  NXC_TRACE(ctx, "nxc_argreduce_exchange");
  nxc_status s = NXC_OK;
  auto fail = [&](nxc_status st) {
    if (st && strcmp(st, NXC_ERR_CUDA) != 0) snprintf(ctx->err, sizeof ctx->err, "%s", st);
    return st;
  };
  if ((s = nxc_check_tensor(out)) || (s = nxc_check_tensor(x)) || (s = nxc_check_tensor(local_idx))) return fail(s);
  if (!ctx->p2p || ctx->dist_poisoned) {
    snprintf(ctx->err, sizeof ctx->err, "%s: nxc_argreduce_exchange needs the peer-memory mailboxes (nxc_dist_p2p_enabled)", NXC_ERR_NCCL);
    return NXC_ERR_NCCL;
  }
  const int dt = x->dtype;
  if (nxc_is_packed(dt)) return fail(NXC_ERR_PACKED);
  if ((nxc_dtype_class(dt) & NXC_CLS_COMPLEX) || out->dtype != NXC_I32 || local_idx->dtype != NXC_I32)
    return fail(NXC_ERR_UNSUPPORTED_DTYPE);
  if (axis < 0 || axis >= x->ndim) return fail(NXC_ERR_AXIS);
  if (out->ndim != x->ndim - 1 || local_idx->ndim != x->ndim - 1) return fail(NXC_ERR_OUT_RANK);
  NxcArgXArgs g;
  g.count = 1;
  g.kept.n = 0;
  int64_t dense = 1;  // out and local_idx must be C-contiguous over the kept shape
  for (int d = x->ndim - 1, j = x->ndim - 2; d >= 0; d--) {
    if (d == axis) continue;
    if (out->shape[j] != x->shape[d] || local_idx->shape[j] != x->shape[d]) return fail(NXC_ERR_SHAPE);
    if (x->shape[d] != 1 && (out->strides[j] != dense || local_idx->strides[j] != dense)) return fail(NXC_ERR_SHAPE);
    dense *= x->shape[d];
    j--;
  }
  for (int d = 0; d < x->ndim; d++) {
    if (d == axis) continue;
    g.kept.shape[g.kept.n] = x->shape[d];
    g.kept.in_stride[g.kept.n] = x->strides[d];
    g.kept.out_stride[g.kept.n] = 0;
    g.count *= x->shape[d];
    g.kept.n++;
  }
  if (x->shape[axis] == 0) return fail(NXC_ERR_EMPTY_REDUCE);
  if (g.count == 0) return NXC_OK;
  if (g.count > nxc_argreduce_exchange_max_outputs(ctx)) return fail(NXC_ERR_TOO_LARGE);
  g.small = g.count < 0x7FFFFFFFLL;
  for (int i = 0; i < g.kept.n; i++) g.kept.div[i] = nxc_fastdiv_make((uint32_t)g.kept.shape[i]);
  g.axis_stride = x->strides[axis];
  g.idx_offset = idx_offset;
  const char *xb = (const char *)x->data + x->offset * nxc_elem_size(dt);
  const int32_t *li = (const int32_t *)local_idx->data + local_idx->offset;
  int32_t *ob = (int32_t *)out->data + out->offset;
  nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
  if (is_max) {
    NXC_DISPATCH_DTYPE(dt, { st = NxcFusedArg<ArgP<1, DT>, ArgP<1, DT>::ok>::go(ctx, g, xb, li, ob); })
  } else {
    NXC_DISPATCH_DTYPE(dt, { st = NxcFusedArg<ArgP<0, DT>, ArgP<0, DT>::ok>::go(ctx, g, xb, li, ob); })
  }
  return fail(st);
}
