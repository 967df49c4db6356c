// nxc_scan.cu -- inclusive cumulative sum / prod / max / min along one axis.
// Replaces caml_nx_c_cum{sum,prod,max,min} (reference: nx_c_fold.c:161-176,
// 834-837; nx_c_engine.c:1293-1331). Slices are independent; within a slice the
// walk is sequential in the compute type and the running value is rounded to
// storage per element, exactly as the reference, so results are bit-identical
// (floats included). One thread per slice; adjacent threads take adjacent slices,
// which coalesces whenever the scanned axis is not the innermost one.
#include "nxc_ops.cuh"
#include "nxc_fold.cuh"

struct ScanArgs {
  NxcDimList kept;
  int64_t slices, len, ai, ao;
  int small;
};

template <int OP, int DT> struct ScanP {
  typedef DT_<DT> D;
  typedef typename D::C C;
  static constexpr int cls = D::cls;
  static constexpr bool ok = (OP == NXC_SUM || OP == NXC_PROD) ? (cls != NXC_CLS_BOOL) : (cls != NXC_CLS_COMPLEX);
  __device__ __forceinline__ static C init() {
    if constexpr (cls == NXC_CLS_COMPLEX) return zmk<C>(OP == NXC_PROD ? 1 : 0, 0);
    else if constexpr (OP == NXC_SUM) return (C)0;
    else if constexpr (OP == NXC_PROD) return (C)1;
    else if constexpr (cls == NXC_CLS_FLOAT) return OP == NXC_RMAX ? (C)-INFINITY : (C)INFINITY;
    else if constexpr (cls == NXC_CLS_BOOL) return OP == NXC_RMAX ? (C)0 : (C)1;
    else if constexpr (cls == NXC_CLS_SINT) {
      if constexpr (sizeof(C) == 8) return OP == NXC_RMAX ? (C)INT64_MIN : (C)INT64_MAX;
      else return OP == NXC_RMAX ? (C)INT32_MIN : (C)INT32_MAX;
    } else return OP == NXC_RMAX ? (C)0 : (C)~(C)0;
  }
  __device__ __forceinline__ static C cmb(C m, C v) {
    if constexpr (cls == NXC_CLS_COMPLEX) return OP == NXC_SUM ? zadd(m, v) : zmul(m, v);
    else if constexpr (OP == NXC_SUM) {
      if constexpr (cls == NXC_CLS_SINT) return (C)((typename UT<C>::U)m + (typename UT<C>::U)v);
      else return m + v;
    } else if constexpr (OP == NXC_PROD) {
      if constexpr (cls == NXC_CLS_SINT) return (C)((typename UT<C>::U)m * (typename UT<C>::U)v);
      else return m * v;
    } else if constexpr (cls == NXC_CLS_FLOAT) {  // sequential NaN-sticky form (nx_c_fold.c:80-89)
      if (OP == NXC_RMAX ? (v > m) : (v < m)) return v;
      if (v != v) return v;
      return m;
    } else return (OP == NXC_RMAX ? (v > m) : (v < m)) ? v : m;
  }
};

template <int OP, int DT>
__global__ void __launch_bounds__(128) scan_kernel(typename DT_<DT>::S *__restrict__ out,
                                                   const typename DT_<DT>::S *__restrict__ in,
                                                   const __grid_constant__ ScanArgs a) {
  typedef DT_<DT> D;
  typedef ScanP<OP, DT> P;
  const int64_t s = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (s >= a.slices) return;
  int64_t io, oo;
  nxc_dims_offset(a.kept, s, a.small, io, oo);
  typename D::C acc = P::init();
  for (int64_t k = 0; k < a.len; k++) {
    acc = P::cmb(acc, D::ld(in[io + k * a.ai]));
    out[oo + k * a.ao] = D::st(acc);
  }
}

template <int OP, int DT, bool OK> struct ScanLaunch {
  static nxc_status go(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in, const ScanArgs &a) {
    typedef typename DT_<DT>::S S;
    const int64_t blocks = (a.slices + 127) / 128;
    scan_kernel<OP, DT><<<(unsigned)blocks, 128, 0, ctx->stream>>>((S *)out->data + out->offset,
                                                                  (const S *)in->data + in->offset, a);
    NXC_LAUNCH_CHECK(ctx);
    return NXC_OK;
  }
};
template <int OP, int DT> struct ScanLaunch<OP, DT, false> {
  static nxc_status go(nxc_ctx *, const nxc_tensor *, const nxc_tensor *, const ScanArgs &) { return NXC_ERR_UNSUPPORTED_DTYPE; }
};

#define SCAN_CASE(OPC) \
  case OPC: { NXC_DISPATCH_DTYPE(dt, { st = ScanLaunch<OPC, DT, ScanP<OPC, DT>::ok>::go(ctx, out, in, a); }) } break;

extern "C" nxc_status nxc_scan(nxc_ctx *ctx, int op, const nxc_tensor *out, const nxc_tensor *in, int axis) {
  NXC_TRACE(ctx, "nxc_scan");
  nxc_status s;
  if ((s = nxc_check_tensor(in)) || (s = nxc_check_tensor(out))) goto fail;
  {
    const int dt = in->dtype, cls = nxc_dtype_class(dt);
    if (op < 0 || op >= NXC_REDUCE_COUNT) { s = NXC_ERR_BAD_OP; goto fail; }
    if (cls & NXC_CLS_PACKED) { s = NXC_ERR_PACKED; goto fail; }
    const bool arith = (op == NXC_SUM || op == NXC_PROD);
    if ((arith && (cls & NXC_CLS_BOOL)) || (!arith && (cls & NXC_CLS_COMPLEX)) || out->dtype != dt) { s = NXC_ERR_UNSUPPORTED_DTYPE; goto fail; }
    if (axis < 0 || axis >= in->ndim) { s = NXC_ERR_AXIS; goto fail; }
    if (out->ndim != in->ndim) { s = NXC_ERR_OUT_RANK; goto fail; }
    ScanArgs a;
    int64_t ks[NXC_MAX_NDIM], ki[NXC_MAX_NDIM], ko[NXC_MAX_NDIM];
    int n = 0;
    a.slices = 1;
    for (int d = 0; d < in->ndim; d++) {
      if (d == axis) continue;
      ks[n] = in->shape[d]; ki[n] = in->strides[d]; ko[n] = out->strides[d]; n++;
      a.slices *= in->shape[d];
    }
    a.len = in->shape[axis];
    a.ai = in->strides[axis];
    a.ao = out->strides[axis];
    if (a.slices == 0 || a.len == 0) return NXC_OK;
    a.small = a.slices < 0x7FFFFFFFLL;
    nxc_dimlist_set(a.kept, n, ks, ki, ko, a.small);
    nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
    switch (op) { SCAN_CASE(NXC_SUM) SCAN_CASE(NXC_PROD) SCAN_CASE(NXC_RMAX) SCAN_CASE(NXC_RMIN) default: break; }
    s = st;
    if (s) goto fail;
    return NXC_OK;
  }
fail:
  if (s && strcmp(s, NXC_ERR_CUDA) != 0) snprintf(ctx->err, sizeof ctx->err, "%s", s);
  return s;
}
