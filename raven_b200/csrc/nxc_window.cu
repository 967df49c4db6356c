// nxc_window.cu -- unfold (im2col) and fold (col2im), the primitives Kaun's conv / pool are
// built on (backend_intf.ml:532-574). Replaces caml_nx_c_unfold / caml_nx_c_fold (reference:
// nx_c_move.c:588-870).
//   unfold  (leading..., spatial...) -> (leading..., prod(kernel), L): every output element
//           copies one input element, or is zero when its tap falls in the padding.
//   fold    (leading..., prod(kernel), L) -> (leading..., output...): parallel over OUTPUT
//           elements; each sums the taps that land on it, walking the kernel offsets in
//           row-major order in the dtype's compute type -- the same order as the reference,
//           so results are bit-identical (floats included) and no atomics are needed.
#include "nxc_ops.cuh"
#include "nxc_fold.cuh"

#define WIN_MAX 8  // spatial dims supported on the device (the reference allows up to 32)

struct WinArgs {
  int ld, K, small;
  NxcDimList lead;  // leading dims: in / out strides
  int64_t kernel[WIN_MAX], stride[WIN_MAX], dilation[WIN_MAX], pad_before[WIN_MAX];
  int64_t win[WIN_MAX], win_cum[WIN_MAX], ker_cum[WIN_MAX], extent[WIN_MAX];
  int64_t sp_stride[WIN_MAX];  // strides of the spatial dims (input of unfold / output of fold)
  int64_t kstride, lstride;    // strides of the (kernel_prod, L) dims of the column tensor
  int64_t kernel_prod, L, total, spatial_total;
};

template <class T>
__global__ void __launch_bounds__(256) unfold_kernel(T *__restrict__ out, const T *__restrict__ in,
                                                     const __grid_constant__ WinArgs a) {
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int64_t it = (int64_t)blockIdx.x * 256 + threadIdx.x; it < a.total; it += step) {
    const int64_t wf = it % a.L, rem = it / a.L, kf = rem % a.kernel_prod, lead = rem / a.kernel_prod;
    int64_t io, oo;
    nxc_dims_offset(a.lead, lead, a.small, io, oo);
    oo += kf * a.kstride + wf * a.lstride;
    bool valid = true;
    for (int d = 0; d < a.K; d++) {
      const int64_t wc = (wf / a.win_cum[d]) % a.win[d], kc = (kf / a.ker_cum[d]) % a.kernel[d];
      const int64_t sp = wc * a.stride[d] + kc * a.dilation[d] - a.pad_before[d];
      if (sp < 0 || sp >= a.extent[d]) { valid = false; break; }
      io += sp * a.sp_stride[d];
    }
    T z;
    memset(&z, 0, sizeof z);
    out[oo] = valid ? in[io] : z;
  }
}

template <int DT>
__global__ void __launch_bounds__(256) fold_kernel(typename DT_<DT>::S *__restrict__ out,
                                                   const typename DT_<DT>::S *__restrict__ in,
                                                   const __grid_constant__ WinArgs a) {
  typedef DT_<DT> D;
  typedef typename D::C C;
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int64_t it = (int64_t)blockIdx.x * 256 + threadIdx.x; it < a.total; it += step) {
    const int64_t o_lin = it % a.spatial_total, lead = it / a.spatial_total;
    int64_t io, oo;
    nxc_dims_offset(a.lead, lead, a.small, io, oo);
    int64_t oc[WIN_MAX];
    {
      int64_t r = o_lin;
      for (int d = a.K - 1; d >= 0; d--) { oc[d] = r % a.extent[d]; r /= a.extent[d]; oo += oc[d] * a.sp_stride[d]; }
    }
    C acc;
    if constexpr (D::cls == NXC_CLS_COMPLEX) acc = zmk<C>(0, 0); else acc = (C)0;
    for (int64_t kf = 0; kf < a.kernel_prod; kf++) {
      bool valid = true;
      int64_t wf = 0;
      for (int d = 0; d < a.K; d++) {
        const int64_t kc = (kf / a.ker_cum[d]) % a.kernel[d];
        const int64_t num = oc[d] + a.pad_before[d] - kc * a.dilation[d];
        if (num < 0 || num % a.stride[d] != 0) { valid = false; break; }
        const int64_t wc = num / a.stride[d];
        if (wc >= a.win[d]) { valid = false; break; }
        wf += wc * a.win_cum[d];
      }
      if (!valid) continue;
      const C v = D::ld(in[io + kf * a.kstride + wf * a.lstride]);
      if constexpr (D::cls == NXC_CLS_COMPLEX) acc = zadd(acc, v);
      else if constexpr (D::cls == NXC_CLS_BOOL) acc = (acc + v) & 0xFFu;  // the reference sums bools in uint8_t
      else if constexpr (D::cls == NXC_CLS_SINT) acc = (C)((typename UT<C>::U)acc + (typename UT<C>::U)v);
      else acc = acc + v;
    }
    out[oo] = D::st(acc);
  }
}

static nxc_status win_setup(WinArgs &a, int K, int ld, const nxc_tensor *lead_in, const nxc_tensor *lead_out,
                            const int64_t *kernel, const int64_t *stride, const int64_t *dilation,
                            const int64_t *padding, const int64_t *extent) {
  if (K < 1 || K > WIN_MAX) return NXC_ERR_SHAPE;
  a.K = K;
  a.ld = ld;
  int64_t lead_total = 1;
  for (int d = 0; d < ld; d++) lead_total *= lead_in->shape[d];
  a.small = lead_total < 0x7FFFFFFFLL;
  nxc_dimlist_set(a.lead, ld, lead_in->shape, lead_in->strides, lead_out->strides, a.small);
  for (int d = 0; d < K; d++) {
    a.kernel[d] = kernel[d]; a.stride[d] = stride[d]; a.dilation[d] = dilation[d];
    a.pad_before[d] = padding[2 * d];
    a.extent[d] = extent[d];
    const int64_t eff = dilation[d] * (kernel[d] - 1) + 1;
    const int64_t padded = extent[d] + padding[2 * d] + padding[2 * d + 1];
    if (stride[d] <= 0) return NXC_ERR_SHAPE;
    const int64_t w = (padded - eff) / stride[d] + 1;
    a.win[d] = w < 1 ? 1 : w;
  }
  a.win_cum[K - 1] = 1;
  a.ker_cum[K - 1] = 1;
  for (int d = K - 2; d >= 0; d--) {
    a.win_cum[d] = a.win_cum[d + 1] * a.win[d + 1];
    a.ker_cum[d] = a.ker_cum[d + 1] * a.kernel[d + 1];
  }
  return NXC_OK;
}
static unsigned win_grid(nxc_ctx *ctx, int64_t total) {
  int64_t b = (total + 255) / 256, cap = (int64_t)ctx->sm_count * 32;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}
static nxc_status win_fail(nxc_ctx *ctx, nxc_status s) {
  if (s && strcmp(s, NXC_ERR_CUDA) != 0) snprintf(ctx->err, sizeof ctx->err, "%s", s);
  return s;
}

extern "C" nxc_status nxc_unfold(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in, int K,
                                 const int64_t *kernel, const int64_t *stride, const int64_t *dilation,
                                 const int64_t *padding_flat) {
  NXC_TRACE(ctx, "nxc_unfold");
  nxc_status s;
  if ((s = nxc_check_tensor(out)) || (s = nxc_check_tensor(in))) return win_fail(ctx, s);
  if (nxc_is_packed(out->dtype)) return win_fail(ctx, NXC_ERR_PACKED);
  if (in->dtype != out->dtype) return win_fail(ctx, NXC_ERR_UNSUPPORTED_DTYPE);
  if (K < 1 || in->ndim < K || out->ndim != in->ndim - K + 2) return win_fail(ctx, NXC_ERR_SHAPE);
  const int ld = in->ndim - K;
  WinArgs a;
  if ((s = win_setup(a, K, ld, in, out, kernel, stride, dilation, padding_flat, &in->shape[ld]))) return win_fail(ctx, s);
  for (int d = 0; d < K; d++) a.sp_stride[d] = in->strides[ld + d];
  a.kernel_prod = out->shape[ld];
  a.L = out->shape[ld + 1];
  a.kstride = out->strides[ld];
  a.lstride = out->strides[ld + 1];
  int64_t lead_total = 1;
  for (int d = 0; d < ld; d++) lead_total *= in->shape[d];
  a.total = lead_total * a.kernel_prod * a.L;
  a.spatial_total = 0;
  if (a.total == 0) return NXC_OK;
  const int64_t es = nxc_elem_size(out->dtype);
  char *ob = (char *)out->data + out->offset * es;
  const char *ib = (const char *)in->data + in->offset * es;
  const unsigned g = win_grid(ctx, a.total);
  switch (es) {
    case 1: unfold_kernel<uint8_t><<<g, 256, 0, ctx->stream>>>((uint8_t *)ob, (const uint8_t *)ib, a); break;
    case 2: unfold_kernel<uint16_t><<<g, 256, 0, ctx->stream>>>((uint16_t *)ob, (const uint16_t *)ib, a); break;
    case 4: unfold_kernel<uint32_t><<<g, 256, 0, ctx->stream>>>((uint32_t *)ob, (const uint32_t *)ib, a); break;
    case 8: unfold_kernel<uint2><<<g, 256, 0, ctx->stream>>>((uint2 *)ob, (const uint2 *)ib, a); break;
    default: unfold_kernel<uint4><<<g, 256, 0, ctx->stream>>>((uint4 *)ob, (const uint4 *)ib, a); break;
  }
  NXC_LAUNCH_CHECK(ctx);
  return NXC_OK;
}

extern "C" nxc_status nxc_fold(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in, int K,
                               const int64_t *output_size, const int64_t *kernel, const int64_t *stride,
                               const int64_t *dilation, const int64_t *padding_flat) {
  NXC_TRACE(ctx, "nxc_fold");
  nxc_status s;
  if ((s = nxc_check_tensor(out)) || (s = nxc_check_tensor(in))) return win_fail(ctx, s);
  const int dt = out->dtype;
  if (nxc_is_packed(dt)) return win_fail(ctx, NXC_ERR_PACKED);
  if (in->dtype != dt) return win_fail(ctx, NXC_ERR_UNSUPPORTED_DTYPE);
  if (K < 1 || in->ndim < 2 || out->ndim != in->ndim - 2 + K) return win_fail(ctx, NXC_ERR_SHAPE);
  const int ld = in->ndim - 2;
  WinArgs a;
  if ((s = win_setup(a, K, ld, in, out, kernel, stride, dilation, padding_flat, output_size))) return win_fail(ctx, s);
  // the leading dims are indexed through OUT's shape in the reference; they agree with in's
  for (int d = 0; d < K; d++) a.sp_stride[d] = out->strides[ld + d];
  a.kernel_prod = in->shape[ld];
  a.L = in->shape[ld + 1];
  a.kstride = in->strides[ld];
  a.lstride = in->strides[ld + 1];
  int64_t lead_total = 1;
  for (int d = 0; d < ld; d++) lead_total *= in->shape[d];
  a.spatial_total = 1;
  for (int d = 0; d < K; d++) a.spatial_total *= output_size[d];
  a.total = lead_total * a.spatial_total;
  if (a.total == 0) return NXC_OK;
  {  // the reference trusts the frontend's column shape; a device read past it would poison the
     // context, so a column tensor that does not match the window geometry is refused
    int64_t nwin = 1, kp = 1;
    for (int d = 0; d < K; d++) { nwin *= a.win[d]; kp *= kernel[d]; }
    if (a.L != nwin || a.kernel_prod != kp) return win_fail(ctx, NXC_ERR_SHAPE);
  }
  nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
  const unsigned g = win_grid(ctx, a.total);
  NXC_DISPATCH_DTYPE(dt, {
    typedef typename DT_<DT>::S S;
    fold_kernel<DT><<<g, 256, 0, ctx->stream>>>((S *)out->data + out->offset, (const S *)in->data + in->offset, a);
    ctx->launches++;
    st = NXC_OK;
  })
  if (st) return win_fail(ctx, st);
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? NXC_OK : nxc_cuda_fail(ctx, e, "fold launch");
}
