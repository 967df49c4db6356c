// nxc_sort.cu -- sort / argsort along one axis.
// Replaces caml_nx_c_sort / caml_nx_c_argsort (reference: nx_c_sort.c). Semantics kept:
// NaN-class elements (any NaN part for complex) go LAST in both directions, in their original
// order; complex orders lexicographically (real, then imaginary); argsort is stable: value ties
// keep the first index first in either direction (nx_c_sort.c:47-101, 265-276). The reference
// sorts each slice with a serial introsort; here every slice is padded to a power of two and
// run through a bitonic network on (key, original index) pairs -- the index tie-break makes
// the order total, so the network's result is unique and equals the reference's.
//   local kernel   1024-element chunks sorted / merged entirely in shared memory
//   global kernel  one compare-exchange stage for partner distances >= 1024
#include "nxc_ops.cuh"
#include "nxc_fold.cuh"

#define SORT_CH 1024
#define SORT_PAD 0x7FFFFFFF

template <int DT> struct SortCmp {
  typedef DT_<DT> D;
  typedef typename D::S S;
  typedef typename D::C C;
  __device__ __forceinline__ static bool isnan_(C x) {
    if constexpr (D::cls == NXC_CLS_FLOAT) return x != x;
    else if constexpr (D::cls == NXC_CLS_COMPLEX) return x.re != x.re || x.im != x.im;
    else return false;
  }
  __device__ __forceinline__ static bool lt(C x, C y) {
    if constexpr (D::cls == NXC_CLS_COMPLEX) return x.re < y.re || (x.re == y.re && x.im < y.im);
    else return x < y;
  }
  __device__ __forceinline__ static bool eq(C x, C y) {
    if constexpr (D::cls == NXC_CLS_COMPLEX) return x.re == y.re && x.im == y.im;
    else return x == y;
  }
  // does (a, ia) sort before (b, ib)?
  __device__ __forceinline__ static bool before(S a, int32_t ia, S b, int32_t ib, int desc) {
    if (ib == SORT_PAD) return ia != SORT_PAD;
    if (ia == SORT_PAD) return false;
    const C x = D::ld(a), y = D::ld(b);
    const bool nx = isnan_(x), ny = isnan_(y);
    if (nx || ny) return (nx && ny) ? (ia < ib) : ny;
    if (eq(x, y)) return ia < ib;
    return desc ? lt(y, x) : lt(x, y);
  }
};

struct SortArgs {
  NxcDimList kept;  // dims other than the sorted axis (in / out strides)
  int64_t slices, L, P, ai, ao;
  int small, desc, is_arg;
};

template <int DT>
__global__ void __launch_bounds__(256) sort_load_kernel(const typename DT_<DT>::S *__restrict__ in,
                                                        typename DT_<DT>::S *__restrict__ keys, int32_t *__restrict__ idx,
                                                        const __grid_constant__ SortArgs a) {
  const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (g >= a.slices * a.P) return;
  const int64_t s = g / a.P, p = g - s * a.P;
  if (p < a.L) {
    int64_t io, oo;
    nxc_dims_offset(a.kept, s, a.small, io, oo);
    keys[g] = in[io + p * a.ai];
    idx[g] = (int32_t)p;
  } else {
    idx[g] = SORT_PAD;
  }
}
template <int DT>
__global__ void __launch_bounds__(256) sort_store_kernel(typename DT_<DT>::S *__restrict__ out_v, int32_t *__restrict__ out_i,
                                                         const typename DT_<DT>::S *__restrict__ keys,
                                                         const int32_t *__restrict__ idx, const __grid_constant__ SortArgs a) {
  const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (g >= a.slices * a.L) return;
  const int64_t s = g / a.L, p = g - s * a.L;
  int64_t io, oo;
  nxc_dims_offset(a.kept, s, a.small, io, oo);
  if (a.is_arg) out_i[oo + p * a.ao] = idx[s * a.P + p];
  else out_v[oo + p * a.ao] = keys[s * a.P + p];
}

// All stages (k, j) with j < SORT_CH for k in [k_lo, k_hi] inside shared memory.
template <int DT>
__global__ void __launch_bounds__(SORT_CH / 2) sort_local_kernel(typename DT_<DT>::S *__restrict__ keys, int32_t *__restrict__ idx,
                                                                 int64_t P, int64_t k_lo, int64_t k_hi, int desc) {
  typedef typename DT_<DT>::S S;
  __shared__ S sk[SORT_CH];
  __shared__ int32_t si[SORT_CH];
  const int64_t base = (int64_t)blockIdx.x * SORT_CH;
  for (int t = threadIdx.x; t < SORT_CH; t += SORT_CH / 2) { sk[t] = keys[base + t]; si[t] = idx[base + t]; }
  __syncthreads();
  for (int64_t k = k_lo; k <= k_hi; k <<= 1) {
    int64_t j0 = (k >> 1) < SORT_CH ? (k >> 1) : (SORT_CH >> 1);
    for (int64_t j = j0; j > 0; j >>= 1) {
      const int t = threadIdx.x;
      const int lo = (int)(((t & ~((int)j - 1)) << 1) | (t & ((int)j - 1)));
      const int hi = lo + (int)j;
      const int64_t p = (base + lo) & (P - 1);
      const bool up = (p & k) == 0;
      const bool b = SortCmp<DT>::before(sk[hi], si[hi], sk[lo], si[lo], desc);
      if (b == up) {
        S tk = sk[lo]; sk[lo] = sk[hi]; sk[hi] = tk;
        int32_t ti = si[lo]; si[lo] = si[hi]; si[hi] = ti;
      }
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < SORT_CH; t += SORT_CH / 2) { keys[base + t] = sk[t]; idx[base + t] = si[t]; }
}
template <int DT>
__global__ void __launch_bounds__(256) sort_global_kernel(typename DT_<DT>::S *__restrict__ keys, int32_t *__restrict__ idx,
                                                          int64_t total_pairs, int64_t P, int64_t k, int64_t j, int desc) {
  typedef typename DT_<DT>::S S;
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= total_pairs) return;
  const int64_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
  const int64_t hi = lo + j;
  const bool up = ((lo & (P - 1)) & k) == 0;
  const S a = keys[lo], b = keys[hi];
  const int32_t ia = idx[lo], ib = idx[hi];
  if (SortCmp<DT>::before(b, ib, a, ia, desc) == up) {
    keys[lo] = b; keys[hi] = a;
    idx[lo] = ib; idx[hi] = ia;
  }
}

template <int DT>
static nxc_status sort_run(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in, SortArgs &a) {
  typedef typename DT_<DT>::S S;
  int64_t P = 1;
  while (P < a.L) P <<= 1;
  if (P < SORT_CH) {
    // several slices per 1024-chunk: the total element count must be a multiple of the chunk
  }
  a.P = P;
  const int64_t total = a.slices * P;
  const int64_t padded = (total + SORT_CH - 1) / SORT_CH * SORT_CH;
  void *scr;
  nxc_status s = nxc_scratch(ctx, (size_t)padded * (sizeof(S) + 4) + 256, &scr);
  if (s) return s;
  S *keys = (S *)scr;
  int32_t *idx = (int32_t *)((char *)scr + (((size_t)padded * sizeof(S) + 255) & ~(size_t)255));
  if (padded > total) NXC_CUDA_TRY(ctx, cudaMemsetAsync(idx + total, 0x7F, (size_t)(padded - total) * 4, ctx->stream));
  const S *ib = (const S *)in->data + in->offset;
  sort_load_kernel<DT><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(ib, keys, idx, a);
  NXC_LAUNCH_CHECK(ctx);
  const unsigned chunks = (unsigned)(padded / SORT_CH);
  const int64_t k_local = P < SORT_CH ? P : SORT_CH;
  if (P >= 2) {
    sort_local_kernel<DT><<<chunks, SORT_CH / 2, 0, ctx->stream>>>(keys, idx, P, 2, k_local, a.desc);
    NXC_LAUNCH_CHECK(ctx);
  }
  for (int64_t k = SORT_CH * 2; k <= P; k <<= 1) {
    for (int64_t j = k >> 1; j >= SORT_CH; j >>= 1) {
      sort_global_kernel<DT><<<(unsigned)((total / 2 + 255) / 256), 256, 0, ctx->stream>>>(keys, idx, total / 2, P, k, j, a.desc);
      NXC_LAUNCH_CHECK(ctx);
    }
    sort_local_kernel<DT><<<chunks, SORT_CH / 2, 0, ctx->stream>>>(keys, idx, P, k, k, a.desc);
    NXC_LAUNCH_CHECK(ctx);
  }
  const int64_t n_out = a.slices * a.L;
  sort_store_kernel<DT><<<(unsigned)((n_out + 255) / 256), 256, 0, ctx->stream>>>(
      a.is_arg ? nullptr : (S *)out->data + out->offset, a.is_arg ? (int32_t *)out->data + out->offset : nullptr, keys, idx, a);
  NXC_LAUNCH_CHECK(ctx);
  return NXC_OK;
}

extern "C" nxc_status nxc_sort(nxc_ctx *ctx, int is_arg, const nxc_tensor *out, const nxc_tensor *in, int axis,
                               int descending) {
  NXC_TRACE(ctx, "nxc_sort");
  nxc_status s;
  if ((s = nxc_check_tensor(in)) || (s = nxc_check_tensor(out))) goto fail;
  {
    const int dt = in->dtype;
    if (nxc_is_packed(dt)) { s = NXC_ERR_PACKED; goto fail; }
    if (out->dtype != (is_arg ? NXC_I32 : dt)) { s = NXC_ERR_UNSUPPORTED_DTYPE; goto fail; }
    if (axis < 0 || axis >= in->ndim) { s = NXC_ERR_AXIS; goto fail; }
    if (out->ndim != in->ndim) { s = NXC_ERR_OUT_RANK; goto fail; }
    SortArgs a;
    int64_t ks[NXC_MAX_NDIM], ki[NXC_MAX_NDIM], ko[NXC_MAX_NDIM];
    int n = 0;
    a.slices = 1;
    for (int d = 0; d < in->ndim; d++) {
      if (out->shape[d] != in->shape[d]) { s = NXC_ERR_SHAPE; goto fail; }
      if (d == axis) continue;
      ks[n] = in->shape[d]; ki[n] = in->strides[d]; ko[n] = out->strides[d]; n++;
      a.slices *= in->shape[d];
    }
    a.L = in->shape[axis];
    if (a.slices == 0 || a.L == 0) return NXC_OK;
    for (int d = 0; d < in->ndim; d++)
      if (in->shape[d] > 1 && out->strides[d] == 0) { s = NXC_ERR_OUT_ALIASED; goto fail; }
    if (a.L > (1LL << 30)) { s = NXC_ERR_ARGREDUCE_CAP; goto fail; }
    a.ai = in->strides[axis];
    a.ao = out->strides[axis];
    a.desc = descending ? 1 : 0;
    a.is_arg = is_arg ? 1 : 0;
    a.small = a.slices < 0x7FFFFFFFLL;
    nxc_dimlist_set(a.kept, n, ks, ki, ko, a.small);
    nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
    NXC_DISPATCH_DTYPE(dt, { st = sort_run<DT>(ctx, out, in, a); })
    s = st;
    if (s) goto fail;
    return NXC_OK;
  }
fail:
  if (s && strcmp(s, NXC_ERR_CUDA) != 0) snprintf(ctx->err, sizeof ctx->err, "%s", s);
  return s;
}
