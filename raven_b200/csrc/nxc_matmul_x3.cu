// nxc_matmul_x3.cu -- f32 matmul on the tcgen05 tensor cores at f32-class accuracy ("3xTF32").
// The reference multiplies f32 in f32 (nx_c_matmul.c:128-311, accumulate in the compute type); a
// plain tf32 product keeps 11 significand bits of each operand (1e-3 relative), which is why tf32
// is opt-in. Splitting every operand element into hi = RN_tf32(a) and lo = a - hi (exact in f32)
// and summing  lo_a*hi_b + hi_a*lo_b + hi_a*hi_b  recovers ~22 bits per product; only lo*lo
// (2^-22 relative) is dropped. That sum is an ORDINARY tf32 GEMM over a concatenated K axis:
//     A' = [A_lo | A_hi | A_hi]  (M x 3K),   B' = [B_hi ; B_lo ; B_hi]  (3K x N),   C = A' B'
// (in the two cross sections a non-finite element's hi part is replaced by 0, see x3_split)
// so the tensor-core kernel is reused unchanged; this file is the split pre-pass (one pass over
// each operand through a 32x32 shared-memory tile, so the source is read along ITS unit-stride dim
// and both A' and B'^T are written K-major, the layout TMA and the UMMA descriptors like best) and
// the plumbing. Small terms are accumulated first.
#include "nxc_matmul.cuh"

namespace {

struct X3Args {
  const float *src;
  float *dst;
  int64_t rows, k, kp;       // kp: k rounded up to the tf32 k-block (32), zero filled
  int64_t rs, ks;            // source element strides along rows / k
  int batch_nd;
  int64_t bshape[NXC_MAX_NDIM], bstr[NXC_MAX_NDIM];
  int lo_section, read_r_fast;  // which of the three K sections holds lo (A: 0, B: 1)
};

// his: the hi part as used in the two CROSS terms (lo*hi, hi*lo). A non-finite element must meet
// its partner exactly once, in the hi*hi section: inf * lo_b would be an infinity with lo_b's
// (arbitrary) sign, or NaN where lo_b == 0, and poison a correctly signed infinity.
__device__ __forceinline__ void x3_split(float a, float &hi, float &his, float &lo) {
  const uint32_t u = __float_as_uint(a);
  if ((u & 0x7F800000u) == 0x7F800000u) { hi = a; his = 0.f; lo = 0.f; return; }  // inf / nan
  uint32_t h = (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;           // RNE to 10 explicit mantissa bits
  if ((h & 0x7F800000u) == 0x7F800000u) h = u & 0xFFFFE000u;           // would round to inf: truncate
  hi = __uint_as_float(h);
  his = hi;
  lo = a - hi;
}

__global__ void __launch_bounds__(256) nxc_x3_split_kernel(const __grid_constant__ X3Args a) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t k0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  int64_t b = blockIdx.z, boff = 0;
  for (int d = a.batch_nd - 1; d >= 0; d--) {
    const int64_t q = b / a.bshape[d];
    boff += (b - q * a.bshape[d]) * a.bstr[d];
    b = q;
  }
  const float *src = a.src + boff;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int y = ty + 8 * i;
    // read_r_fast: lanes walk rows (the source's unit-stride dim), else lanes walk k
    const int64_t r = r0 + (a.read_r_fast ? tx : y), k = k0 + (a.read_r_fast ? y : tx);
    float v = 0.f;
    if (r < a.rows && k < a.k) v = src[r * a.rs + k * a.ks];
    if (a.read_r_fast) tile[tx][y] = v; else tile[y][tx] = v;
  }
  __syncthreads();
  float *dst = a.dst + (int64_t)blockIdx.z * a.rows * 3 * a.kp;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int y = ty + 8 * i;
    const int64_t r = r0 + y, k = k0 + tx;
    if (r >= a.rows) continue;  // k < kp always: kp is a multiple of 32
    float hi, his, lo;
    x3_split(tile[y][tx], hi, his, lo);
    float *row = dst + r * 3 * a.kp + k;
#pragma unroll
    for (int s = 0; s < 3; s++) row[s * a.kp] = (s == a.lo_section) ? lo : (s == 2 ? hi : his);
  }
}

}  // namespace

nxc_status nxc_matmul_f32x3(nxc_ctx *ctx, const NxcMatmulProblem &q) {
  if (q.dt != NXC_F32 || q.k == 0 || q.c_cs != 1) return NXC_MM_TC_DECLINED;
  if (q.nbatch > 65535) return NXC_MM_TC_DECLINED;
  const int64_t kp = (q.k + 31) / 32 * 32;
  if (3 * kp >= 0x7FFFFFFFLL) return NXC_MM_TC_DECLINED;
  bool a_b = false, b_b = false;
  for (int i = 0; i < q.batch_nd; i++) {
    if (q.bshape[i] > 1 && q.as_[i] != 0) a_b = true;
    if (q.bshape[i] > 1 && q.bs_[i] != 0) b_b = true;
  }
  const int64_t nba = a_b ? q.nbatch : 1, nbb = b_b ? q.nbatch : 1;
  float *a3 = NULL, *b3 = NULL;
  nxc_status s = nxc_alloc(ctx, (size_t)(nba * q.m * 3 * kp) * sizeof(float), (void **)&a3);
  if (s) return s;
  if ((s = nxc_alloc(ctx, (size_t)(nbb * q.n * 3 * kp) * sizeof(float), (void **)&b3))) { nxc_free(ctx, a3); return s; }
  auto split = [&](const char *src, float *dst, int64_t rows, int64_t rs, int64_t ks, const int64_t *bstr, bool batched,
                   int lo_section) {
    X3Args x;
    x.src = (const float *)src; x.dst = dst; x.rows = rows; x.k = q.k; x.kp = kp; x.rs = rs; x.ks = ks;
    x.batch_nd = batched ? q.batch_nd : 0;
    for (int i = 0; i < x.batch_nd; i++) { x.bshape[i] = q.bshape[i]; x.bstr[i] = bstr[i]; }
    x.lo_section = lo_section;
    x.read_r_fast = (ks != 1 && rs == 1) ? 1 : 0;
    const dim3 grid((unsigned)(kp / 32), (unsigned)((rows + 31) / 32), (unsigned)(batched ? q.nbatch : 1));
    nxc_x3_split_kernel<<<grid, 256, 0, ctx->stream>>>(x);
    ctx->launches++;
  };
  if ((q.m + 31) / 32 > 65535 || (q.n + 31) / 32 > 65535) { nxc_free(ctx, a3); nxc_free(ctx, b3); return NXC_MM_TC_DECLINED; }
  split(q.a, a3, q.m, q.a_rs, q.a_cs, q.as_, a_b, 0);   // A' = [lo | hi | hi]
  split(q.b, b3, q.n, q.b_cs, q.b_rs, q.bs_, b_b, 1);   // B'^T = [hi | lo | hi]
  if (cudaPeekAtLastError() != cudaSuccess) {
    nxc_free(ctx, a3); nxc_free(ctx, b3);
    return nxc_cuda_fail(ctx, cudaGetLastError(), "matmul");
  }
  NxcMatmulProblem p = q;
  p.k = 3 * kp;
  p.a = (const char *)a3; p.a_rs = 3 * kp; p.a_cs = 1;
  p.b = (const char *)b3; p.b_rs = 1; p.b_cs = 3 * kp;
  int64_t ext = 1;
  for (int i = q.batch_nd - 1; i >= 0; i--) {   // dense batches in C order (or a full broadcast)
    p.as_[i] = (a_b && q.bshape[i] > 1) ? ext * q.m * 3 * kp : 0;
    p.bs_[i] = (b_b && q.bshape[i] > 1) ? ext * q.n * 3 * kp : 0;
    ext *= q.bshape[i];
  }
  // A short, wide-K product leaves most SM pairs idle and is one long dependent K loop (a 256 x 768 x 768 linear
  // of the GPT-2 step at its reference batch: 12 tile pairs, 36 k-blocks each). Its three sections are three
  // independent products lo*hi, hi*lo, hi*hi: run them as three BATCHES of K = kp into a workspace and add the
  // partials in that fixed order (the two cross terms first), one small fold kernel. Same arithmetic per section,
  // a third of the chain, three times the CTAs.
  const int64_t max_ctas = ((q.m + 127) / 128) * ((q.n + 63) / 64);
  if (q.nbatch == 1 && kp >= 256 && max_ctas * 3 <= ctx->sm_count && !getenv("NX_CUDA_X3_NO_SPLIT")) {
    float *ws = NULL;
    if ((s = nxc_alloc(ctx, (size_t)(3 * q.m * q.n) * sizeof(float), (void **)&ws))) { nxc_free(ctx, a3); nxc_free(ctx, b3); return s; }
    NxcMatmulProblem ps = p;
    ps.k = kp;
    ps.batch_nd = 1; ps.nbatch = 3; ps.bshape[0] = 3;
    ps.as_[0] = kp; ps.bs_[0] = kp; ps.cs_[0] = q.m * q.n;
    ps.c = (char *)ws; ps.c_rs = q.n; ps.c_cs = 1;
    s = nxc_matmul_tc(ctx, ps);
    if (!s) {
      nxc_tensor wd, cd;
      memset(&wd, 0, sizeof wd);
      memset(&cd, 0, sizeof cd);
      wd.data = ws; wd.dtype = NXC_F32; wd.ndim = 3;
      wd.shape[0] = 3; wd.shape[1] = q.m; wd.shape[2] = q.n;
      wd.strides[0] = q.m * q.n; wd.strides[1] = q.n; wd.strides[2] = 1;
      cd.data = (void *)q.c; cd.dtype = NXC_F32; cd.ndim = 2;
      cd.shape[0] = q.m; cd.shape[1] = q.n;
      cd.strides[0] = q.c_rs; cd.strides[1] = q.c_cs;
      const int axis0 = 0;
      s = nxc_reduce(ctx, NXC_SUM, &cd, &wd, &axis0, 1);
    }
    nxc_free(ctx, ws);
    nxc_free(ctx, a3);
    nxc_free(ctx, b3);
    return s;
  }
  s = nxc_matmul_tc(ctx, p);
  nxc_free(ctx, a3);   // stream-ordered: reused only after the GEMM that reads them
  nxc_free(ctx, b3);
  return s;
}
