// nxc_packed.cu -- the storage-only int4 / uint4 dtypes: two elements per byte, element i
// of a packed operand is nibble (offset + i): byte (offset+i)>>1, low nibble on even, high
// on odd; int4 sign-extends (reference: nx_c_map.c:1046-1177, nx_c_move.c:151-189).
//
// Packed operands must be dense (contiguous ignoring size-1 dims); anything else is
// "packed dtype not supported for this operation", as in the reference. To keep nibble
// writes race-free every thread owns one destination BYTE (both of its nibbles); the two
// edge bytes of a range merge with the neighbour nibble already stored there, which is how
// an assign into an odd-length prefix preserves the element next to it.
#include "nxc_ops.cuh"
#include "nxc_cast.cuh"

static bool packed_dense(const nxc_tensor *a) {
  int64_t expect = 1;
  for (int i = a->ndim - 1; i >= 0; i--) {
    if (a->shape[i] == 1) continue;
    if (a->strides[i] != expect) return false;
    expect *= a->shape[i];
  }
  return true;
}

__device__ __forceinline__ int f2i4(double v, bool sgn) {
  if (v != v) return 0;
  if (sgn) { if (v <= -8.0) return -8; if (v >= 7.0) return 7; }
  else { if (v <= 0.0) return 0; if (v >= 15.0) return 15; }
  return (int)v;
}

// compute src -> nibble value (wraps for int/bool, saturates for float/complex)
template <int SRC> __device__ __forceinline__ uint32_t to_nibble(typename DT_<SRC>::S s, bool sgn) {
  typedef DT_<SRC> A;
  typename A::C v = A::ld(s);
  if constexpr (A::cls == NXC_CLS_FLOAT) return (uint32_t)f2i4((double)v, sgn) & 0xFu;
  else if constexpr (A::cls == NXC_CLS_COMPLEX) return (uint32_t)f2i4((double)v.re, sgn) & 0xFu;
  else return (uint32_t)v & 0xFu;
}
// nibble value (already sign/zero-extended int) -> dst storage, by the cast policy
template <int DST> __device__ __forceinline__ typename DT_<DST>::S from_nibble(int v) {
  typedef DT_<DST> B;
  typedef typename B::C CB;
  if constexpr (B::cls == NXC_CLS_FLOAT) return B::st((CB)v);
  else if constexpr (B::cls == NXC_CLS_COMPLEX) return zmk<CB>((typename ZR<CB>::R)v, 0);
  else if constexpr (B::cls == NXC_CLS_BOOL) return bool_s{(uint8_t)(v != 0)};
  else return (typename B::S)(CB)v;
}

template <int SRC>
__global__ void __launch_bounds__(256) castp_to_kernel(uint8_t *db, int64_t doff, const typename DT_<SRC>::S *src,
                                                       int64_t n, int sgn) {
  const int64_t first_byte = doff >> 1, last_byte = (doff + n - 1) >> 1;
  const int64_t byte = first_byte + (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (byte > last_byte) return;
  uint8_t cur = db[byte];
  for (int h = 0; h < 2; h++) {
    const int64_t di = byte * 2 + h, i = di - doff;
    if (i < 0 || i >= n) continue;
    const uint32_t nib = to_nibble<SRC>(src[i], sgn != 0);
    cur = h ? (uint8_t)((cur & 0x0F) | (nib << 4)) : (uint8_t)((cur & 0xF0) | nib);
  }
  db[byte] = cur;
}
template <int DST>
__global__ void __launch_bounds__(256) castp_from_kernel(typename DT_<DST>::S *dst, const uint8_t *sb, int64_t soff,
                                                         int64_t n, int sgn) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int64_t si = soff + i;
  const uint8_t by = sb[si >> 1];
  int v;
  if (sgn) v = (si & 1) ? ((int8_t)by >> 4) : ((int8_t)((by & 0x0F) << 4) >> 4);
  else v = (si & 1) ? (by >> 4) : (by & 0x0F);
  dst[i] = from_nibble<DST>(v);
}
__global__ void __launch_bounds__(256) castp_pp_kernel(uint8_t *db, int64_t doff, const uint8_t *sb, int64_t soff,
                                                       int64_t n) {
  const int64_t first_byte = doff >> 1, last_byte = (doff + n - 1) >> 1;
  const int64_t byte = first_byte + (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (byte > last_byte) return;
  uint8_t cur = db[byte];
  for (int h = 0; h < 2; h++) {
    const int64_t di = byte * 2 + h, i = di - doff;
    if (i < 0 || i >= n) continue;
    const int64_t si = soff + i;
    const uint32_t nib = (sb[si >> 1] >> ((si & 1) * 4)) & 0xFu;
    cur = h ? (uint8_t)((cur & 0x0F) | (nib << 4)) : (uint8_t)((cur & 0xF0) | nib);
  }
  db[byte] = cur;
}

nxc_status nxc_cast_packed(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in) {
  if (!packed_dense(out) || !packed_dense(in)) return NXC_ERR_PACKED;
  const int64_t n = nxc_numel(out);
  if (n == 0) return NXC_OK;
  const int src = in->dtype, dst = out->dtype;
  const bool sp = nxc_is_packed(src), dp = nxc_is_packed(dst);
  if (sp && dp) {
    const int64_t nbytes = ((out->offset + n - 1) >> 1) - (out->offset >> 1) + 1;
    castp_pp_kernel<<<(unsigned)((nbytes + 255) / 256), 256, 0, ctx->stream>>>((uint8_t *)out->data, out->offset,
                                                                              (const uint8_t *)in->data, in->offset, n);
    NXC_LAUNCH_CHECK(ctx);
    return NXC_OK;
  }
  nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
  if (dp) {
    const int64_t nbytes = ((out->offset + n - 1) >> 1) - (out->offset >> 1) + 1;
    const unsigned grid = (unsigned)((nbytes + 255) / 256);
    NXC_DISPATCH_DTYPE(src, {
      typedef typename DT_<DT>::S S;
      castp_to_kernel<DT><<<grid, 256, 0, ctx->stream>>>((uint8_t *)out->data, out->offset,
                                                         (const S *)in->data + in->offset, n, dst == NXC_I4);
      ctx->launches++;
      st = NXC_OK;
    })
  } else {
    const unsigned grid = (unsigned)((n + 255) / 256);
    NXC_DISPATCH_DTYPE(dst, {
      typedef typename DT_<DT>::S S;
      castp_from_kernel<DT><<<grid, 256, 0, ctx->stream>>>((S *)out->data + out->offset, (const uint8_t *)in->data,
                                                           in->offset, n, src == NXC_I4);
      ctx->launches++;
      st = NXC_OK;
    })
  }
  if (st) return st;
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? NXC_OK : nxc_cuda_fail(ctx, e, "packed cast launch");
}

// copy / contiguous / assign for packed dtypes: both sides contiguous at offset 0, same
// element count; whole bytes move, an odd tail merges only the low nibble
// (reference: nx_c_move.c:151-189).
nxc_status nxc_copy_packed(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *in) {
  const int64_t n = nxc_numel(out);
  if (n != nxc_numel(in) || out->offset != 0 || in->offset != 0 || !packed_dense(out) || !packed_dense(in))
    return NXC_ERR_PACKED;
  if (n == 0) return NXC_OK;
  castp_pp_kernel<<<(unsigned)(((n + 1) / 2 + 255) / 256), 256, 0, ctx->stream>>>((uint8_t *)out->data, 0,
                                                                                (const uint8_t *)in->data, 0, n);
  NXC_LAUNCH_CHECK(ctx);
  return NXC_OK;
}
