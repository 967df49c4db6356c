// nxc_random.cu -- threefry2x32-20, bit-exact with the reference
// (nx_c_random.c:44-61; Random123 rotation schedule 13,15,26,6,17,29,16,24, key
// injection after every 4th round, parity word 0x1BD11BDA). key / ctr / out are
// int32 tensors of one shape whose last axis has extent 2; one thread per vector.
#include "nxc_map.cuh"

#define NXC_ERR_THREEFRY_SHAPE "threefry: last axis must have extent 2"

struct TfArgs {
  int nd;  // prefix dims
  int small;
  int64_t total;
  NxcFastDiv div[NXC_MAX_NDIM];
  int64_t shape[NXC_MAX_NDIM], sk[NXC_MAX_NDIM], sc[NXC_MAX_NDIM], so[NXC_MAX_NDIM];
  int64_t k_last, c_last, o_last;
};

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__global__ void __launch_bounds__(256) threefry_kernel(int32_t *__restrict__ out, const int32_t *__restrict__ key,
                                                       const int32_t *__restrict__ ctr,
                                                       const __grid_constant__ TfArgs a) {
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int64_t it = (int64_t)blockIdx.x * 256 + threadIdx.x; it < a.total; it += step) {
    int64_t kb = 0, cb = 0, ob = 0, r = it;
    for (int d = a.nd - 1; d >= 0; d--) {
      int64_t q = r / a.shape[d], c = r - q * a.shape[d];
      r = q;
      kb += c * a.sk[d]; cb += c * a.sc[d]; ob += c * a.so[d];
    }
    const uint32_t k0 = (uint32_t)key[kb], k1 = (uint32_t)key[kb + a.k_last];
    const uint32_t ks[3] = {k0, k1, 0x1BD11BDAu ^ k0 ^ k1};
    uint32_t x0 = (uint32_t)ctr[cb] + k0, x1 = (uint32_t)ctr[cb + a.c_last] + k1;
    const int R[8] = {13, 15, 26, 6, 17, 29, 16, 24};
#pragma unroll
    for (int rd = 0; rd < 20; rd++) {
      x0 += x1;
      x1 = rotl32(x1, R[rd % 8]);
      x1 ^= x0;
      if ((rd + 1) % 4 == 0) {
        const int s = (rd + 1) / 4;
        x0 += ks[s % 3];
        x1 += ks[(s + 1) % 3] + (uint32_t)s;
      }
    }
    out[ob] = (int32_t)x0;
    out[ob + a.o_last] = (int32_t)x1;
  }
}

extern "C" nxc_status nxc_threefry(nxc_ctx *ctx, const nxc_tensor *out, const nxc_tensor *key, const nxc_tensor *ctr) {
  NXC_TRACE(ctx, "nxc_threefry");
  nxc_status s;
  if ((s = nxc_check_tensor(out)) || (s = nxc_check_tensor(key)) || (s = nxc_check_tensor(ctr))) goto fail;
  if (out->dtype != NXC_I32 || key->dtype != NXC_I32 || ctr->dtype != NXC_I32) { s = NXC_ERR_UNSUPPORTED_DTYPE; goto fail; }
  if (key->ndim < 1 || key->ndim != ctr->ndim || key->ndim != out->ndim) { s = NXC_ERR_THREEFRY_SHAPE; goto fail; }
  for (int d = 0; d < key->ndim; d++)
    if (key->shape[d] != ctr->shape[d] || key->shape[d] != out->shape[d]) { s = NXC_ERR_THREEFRY_SHAPE; goto fail; }
  if (key->shape[key->ndim - 1] != 2) { s = NXC_ERR_THREEFRY_SHAPE; goto fail; }
  {
    TfArgs a;
    const int last = key->ndim - 1;
    a.nd = last;
    a.total = 1;
    for (int d = 0; d < last; d++) {
      a.shape[d] = key->shape[d]; a.sk[d] = key->strides[d]; a.sc[d] = ctr->strides[d]; a.so[d] = out->strides[d];
      a.total *= key->shape[d];
    }
    a.k_last = key->strides[last]; a.c_last = ctr->strides[last]; a.o_last = out->strides[last];
    if (a.total == 0) return NXC_OK;
    int64_t b = (a.total + 255) / 256, cap = (int64_t)ctx->sm_count * 32;
    threefry_kernel<<<(unsigned)(b < cap ? b : cap), 256, 0, ctx->stream>>>(
        (int32_t *)out->data + out->offset, (const int32_t *)key->data + key->offset,
        (const int32_t *)ctr->data + ctr->offset, a);
    NXC_LAUNCH_CHECK(ctx);
    return NXC_OK;
  }
fail:
  if (s && strcmp(s, NXC_ERR_CUDA) != 0) snprintf(ctx->err, sizeof ctx->err, "%s", s);
  return s;
}
