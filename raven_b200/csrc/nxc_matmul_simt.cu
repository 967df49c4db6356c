// nxc_matmul_simt.cu -- CUDA-core GEMM for every compute dtype at arbitrary
// strides. This is the exact path: f32 accumulates in f32 FFMA, f64 in DFMA,
// integers modulo 2^w, complex natively, f16/bf16/fp8 in f32 -- the reference's
// accumulate-in-compute-type, store-once rule (nx_c_matmul.c:12-28, 363-423).
// Transposed / broadcast views are read through their strides, never
// materialised (reference: nx_c_matmul.c:14-17).
//
// 64x64 output tile per 256-thread block, 4x4 per thread, K staged 16 deep in
// shared memory. The thread->element map of the staging loads follows whichever
// stride of the operand is 1, so both row-major and transposed views coalesce.
#include "nxc_ops.cuh"
#include "nxc_matmul.cuh"
#include "nxc_map.cuh"

#define MM_BM 64
#define MM_BN 64
#define MM_BK 16

struct MmBatch {
  int nd;
  NxcFastDiv div[NXC_MAX_NDIM];
  int64_t as_[NXC_MAX_NDIM], bs_[NXC_MAX_NDIM], cs_[NXC_MAX_NDIM];
};

template <class C> struct MmAcc {
  __device__ __forceinline__ static C zero() { return (C)0; }
  __device__ __forceinline__ static C fma(C a, C b, C c) { return a * b + c; }
};
template <> struct MmAcc<int32_t> {
  __device__ __forceinline__ static int32_t zero() { return 0; }
  __device__ __forceinline__ static int32_t fma(int32_t a, int32_t b, int32_t c) { return (int32_t)((uint32_t)a * (uint32_t)b + (uint32_t)c); }
};
template <> struct MmAcc<int64_t> {
  __device__ __forceinline__ static int64_t zero() { return 0; }
  __device__ __forceinline__ static int64_t fma(int64_t a, int64_t b, int64_t c) { return (int64_t)((uint64_t)a * (uint64_t)b + (uint64_t)c); }
};
template <> struct MmAcc<cf32> {
  __device__ __forceinline__ static cf32 zero() { return zmk<cf32>(0, 0); }
  __device__ __forceinline__ static cf32 fma(cf32 a, cf32 b, cf32 c) { return zadd(c, zmul(a, b)); }
};
template <> struct MmAcc<cf64> {
  __device__ __forceinline__ static cf64 zero() { return zmk<cf64>(0, 0); }
  __device__ __forceinline__ static cf64 fma(cf64 a, cf64 b, cf64 c) { return zadd(c, zmul(a, b)); }
};

template <int DT>
__global__ void __launch_bounds__(256)
nxc_mm_simt_kernel(const typename DT_<DT>::S *__restrict__ A, const typename DT_<DT>::S *__restrict__ B,
                   typename DT_<DT>::S *__restrict__ Cc, int64_t m, int64_t n, int64_t k,
                   int64_t a_rs, int64_t a_cs, int64_t b_rs, int64_t b_cs, int64_t c_rs, int64_t c_cs,
                   int64_t nbatch, const __grid_constant__ MmBatch bt) {
  typedef DT_<DT> D;
  typedef typename D::S S;
  typedef typename D::C C;
  __shared__ C As[MM_BK][MM_BM + 1];
  __shared__ C Bs[MM_BK][MM_BN + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * MM_BM, n0 = (int64_t)blockIdx.x * MM_BN;
  for (int64_t batch = blockIdx.z; batch < nbatch; batch += gridDim.z) {
    int64_t ao = 0, bo = 0, co = 0;
    {
      uint32_t r = (uint32_t)batch;
      for (int i = bt.nd - 1; i >= 0; i--) {
        uint32_t q = nxc_fastdiv(r, bt.div[i]);
        uint32_t c = r - q * bt.div[i].d;
        ao += (int64_t)c * bt.as_[i]; bo += (int64_t)c * bt.bs_[i]; co += (int64_t)c * bt.cs_[i];
        r = q;
      }
    }
    const S *Ab = A + ao;
    const S *Bb = B + bo;
    S *Cb = Cc + co;
    C acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j] = MmAcc<C>::zero();
    for (int64_t k0 = 0; k0 < k; k0 += MM_BK) {
      // stage A[m0:m0+64, k0:k0+16] and B[k0:k0+16, n0:n0+64]
#pragma unroll
      for (int i = 0; i < (MM_BM * MM_BK) / 256; i++) {
        const int e = tid + i * 256;
        int mm, kk;
        if (a_cs == 1) { mm = e / MM_BK; kk = e % MM_BK; } else { kk = e / MM_BM; mm = e % MM_BM; }
        const int64_t gm = m0 + mm, gk = k0 + kk;
        As[kk][mm] = (gm < m && gk < k) ? D::ld(Ab[gm * a_rs + gk * a_cs]) : MmAcc<C>::zero();
      }
#pragma unroll
      for (int i = 0; i < (MM_BN * MM_BK) / 256; i++) {
        const int e = tid + i * 256;
        int nn, kk;
        if (b_rs == 1) { nn = e / MM_BK; kk = e % MM_BK; } else { kk = e / MM_BN; nn = e % MM_BN; }
        const int64_t gn = n0 + nn, gk = k0 + kk;
        Bs[kk][nn] = (gn < n && gk < k) ? D::ld(Bb[gk * b_rs + gn * b_cs]) : MmAcc<C>::zero();
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < MM_BK; kk++) {
        C av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) av[i] = As[kk][ty + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; j++) bv[j] = Bs[kk][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = MmAcc<C>::fma(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int64_t gm = m0 + ty + 16 * i;
      if (gm >= m) continue;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int64_t gn = n0 + tx + 16 * j;
        if (gn < n) Cb[gm * c_rs + gn * c_cs] = D::st(acc[i][j]);
      }
    }
  }
}


// ---- register-tiled kernel for exact f32 / f64 -----------------------------------------
// 128x128 output tile per 256-thread CTA, 8x8 accumulators per thread (two 4-wide strips in
// each direction so shared-memory reads are 128-bit and conflict-free), K staged 8 deep,
// global -> register prefetch of the next K-slab while the current one is multiplied, one
// barrier per slab. The staging map follows whichever stride of the operand is 1 (128-bit
// loads when rows are 16-byte aligned), so Rune's transposed views (reverse.ml:585-654) stream
// as fast as row-major ones. FFMA/DFMA roof, not the tensor roof: this is the EXACT f32 path.
#define MB_M 128
#define MB_N 128
#define MB_K 8

template <class T>
__global__ void __launch_bounds__(256)
nxc_mm_simt128_kernel(const T *__restrict__ A, const T *__restrict__ B, T *__restrict__ Cc, int64_t m, int64_t n,
                      int64_t k, int64_t a_rs, int64_t a_cs, int64_t b_rs, int64_t b_cs, int64_t c_rs, int64_t c_cs,
                      int64_t nbatch, int a_vec, int b_vec, const __grid_constant__ MmBatch bt) {
  __shared__ __align__(16) T As[2][MB_K][MB_M];
  __shared__ __align__(16) T Bs[2][MB_K][MB_N];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * MB_M, n0 = (int64_t)blockIdx.x * MB_N;
  for (int64_t batch = blockIdx.z; batch < nbatch; batch += gridDim.z) {
    int64_t ao = 0, bo = 0, co = 0;
    {
      uint32_t r = (uint32_t)batch;
      for (int i = bt.nd - 1; i >= 0; i--) {
        uint32_t q = nxc_fastdiv(r, bt.div[i]);
        uint32_t c = r - q * bt.div[i].d;
        ao += (int64_t)c * bt.as_[i]; bo += (int64_t)c * bt.bs_[i]; co += (int64_t)c * bt.cs_[i];
        r = q;
      }
    }
    const T *Ab = A + ao;
    const T *Bb = B + bo;
    T *Cb = Cc + co;
    // staging coordinates: 1024 elements per operand per slab = 4 per thread
    // a_vec == 1: K contiguous  -> thread = (row t/2, 4 consecutive k)
    // a_vec == 2: M contiguous  -> thread = (k t/32, 4 consecutive rows)
    // a_vec == 0: generic       -> 4 scalar loads, (k = e / 128, row = e % 128)
    T ra[4], rb[4];
    auto load_a = [&](int64_t k0) {
      if (a_vec == 1) {
        const int row = tid >> 1, kq = (tid & 1) * 4;
        const int64_t gm = m0 + row;
        if (gm < m && k0 + kq + 3 < k) {
          const float4 *p4 = nullptr; (void)p4;
          const T *p = Ab + gm * a_rs + (k0 + kq);
#pragma unroll
          for (int i = 0; i < 4; i++) ra[i] = p[i];
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) ra[i] = (gm < m && k0 + kq + i < k) ? Ab[gm * a_rs + (k0 + kq + i) * a_cs] : (T)0;
        }
      } else if (a_vec == 2) {
        const int kk = tid >> 5, mq = (tid & 31) * 4;
        const int64_t gk = k0 + kk;
#pragma unroll
        for (int i = 0; i < 4; i++) ra[i] = (gk < k && m0 + mq + i < m) ? Ab[(m0 + mq + i) * a_rs + gk * a_cs] : (T)0;
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int e = tid + i * 256, kk = e >> 7, row = e & 127;
          ra[i] = (m0 + row < m && k0 + kk < k) ? Ab[(m0 + row) * a_rs + (k0 + kk) * a_cs] : (T)0;
        }
      }
    };
    auto store_a = [&](int buf) {
      if (a_vec == 1) {
        const int row = tid >> 1, kq = (tid & 1) * 4;
#pragma unroll
        for (int i = 0; i < 4; i++) As[buf][kq + i][row] = ra[i];
      } else if (a_vec == 2) {
        const int kk = tid >> 5, mq = (tid & 31) * 4;
#pragma unroll
        for (int i = 0; i < 4; i++) As[buf][kk][mq + i] = ra[i];
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++) { const int e = tid + i * 256; As[buf][e >> 7][e & 127] = ra[i]; }
      }
    };
    auto load_b = [&](int64_t k0) {
      if (b_vec == 1) {  // N contiguous
        const int kk = tid >> 5, nq = (tid & 31) * 4;
        const int64_t gk = k0 + kk;
#pragma unroll
        for (int i = 0; i < 4; i++) rb[i] = (gk < k && n0 + nq + i < n) ? Bb[gk * b_rs + (n0 + nq + i) * b_cs] : (T)0;
      } else if (b_vec == 2) {  // K contiguous
        const int col = tid >> 1, kq = (tid & 1) * 4;
        const int64_t gn = n0 + col;
#pragma unroll
        for (int i = 0; i < 4; i++) rb[i] = (gn < n && k0 + kq + i < k) ? Bb[(k0 + kq + i) * b_rs + gn * b_cs] : (T)0;
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int e = tid + i * 256, kk = e >> 7, col = e & 127;
          rb[i] = (n0 + col < n && k0 + kk < k) ? Bb[(k0 + kk) * b_rs + (n0 + col) * b_cs] : (T)0;
        }
      }
    };
    auto store_b = [&](int buf) {
      if (b_vec == 1) {
        const int kk = tid >> 5, nq = (tid & 31) * 4;
#pragma unroll
        for (int i = 0; i < 4; i++) Bs[buf][kk][nq + i] = rb[i];
      } else if (b_vec == 2) {
        const int col = tid >> 1, kq = (tid & 1) * 4;
#pragma unroll
        for (int i = 0; i < 4; i++) Bs[buf][kq + i][col] = rb[i];
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++) { const int e = tid + i * 256; Bs[buf][e >> 7][e & 127] = rb[i]; }
      }
    };

    T acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[i][j] = (T)0;

    const int64_t nslab = (k + MB_K - 1) / MB_K;
    if (nslab > 0) {
      load_a(0); load_b(0);
      store_a(0); store_b(0);
    }
    __syncthreads();
    for (int64_t sidx = 0; sidx < nslab; sidx++) {
      const int buf = (int)(sidx & 1);
      if (sidx + 1 < nslab) { load_a((sidx + 1) * MB_K); load_b((sidx + 1) * MB_K); }
#pragma unroll
      for (int kk = 0; kk < MB_K; kk++) {
        T av[8], bv[8];
#pragma unroll
        for (int i = 0; i < 4; i++) { av[i] = As[buf][kk][ty * 4 + i]; av[4 + i] = As[buf][kk][64 + ty * 4 + i]; }
#pragma unroll
        for (int j = 0; j < 4; j++) { bv[j] = Bs[buf][kk][tx * 4 + j]; bv[4 + j] = Bs[buf][kk][64 + tx * 4 + j]; }
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 8; j++) acc[i][j] = av[i] * bv[j] + acc[i][j];
      }
      if (sidx + 1 < nslab) { store_a(buf ^ 1); store_b(buf ^ 1); }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (gm >= m) continue;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int64_t gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        if (gn < n) Cb[gm * c_rs + gn * c_cs] = acc[i][j];
      }
    }
    __syncthreads();
  }
}

template <class T>
static nxc_status launch_big(nxc_ctx *ctx, const NxcMatmulProblem &p, const MmBatch &bt) {
  dim3 grid((unsigned)((p.n + MB_N - 1) / MB_N), (unsigned)((p.m + MB_M - 1) / MB_M),
            (unsigned)(p.nbatch < 65535 ? p.nbatch : 65535));
  if (grid.y > 65535) return NXC_ERR_SHAPE;
  const int a_vec = p.a_cs == 1 ? 1 : (p.a_rs == 1 ? 2 : 0);
  const int b_vec = p.b_cs == 1 ? 1 : (p.b_rs == 1 ? 2 : 0);
  nxc_mm_simt128_kernel<T><<<grid, 256, 0, ctx->stream>>>((const T *)p.a, (const T *)p.b, (T *)p.c, p.m, p.n, p.k,
                                                         p.a_rs, p.a_cs, p.b_rs, p.b_cs, p.c_rs, p.c_cs, p.nbatch,
                                                         a_vec, b_vec, bt);
  NXC_LAUNCH_CHECK(ctx);
  return NXC_OK;
}

template <int DT, bool OK> struct MmLaunch {
  static nxc_status go(nxc_ctx *ctx, const NxcMatmulProblem &p, dim3 grid, const MmBatch &bt) {
    typedef typename DT_<DT>::S S;
    nxc_mm_simt_kernel<DT><<<grid, 256, 0, ctx->stream>>>((const S *)p.a, (const S *)p.b, (S *)p.c, p.m, p.n,
                                                         p.k, p.a_rs, p.a_cs, p.b_rs, p.b_cs, p.c_rs,
                                                         p.c_cs, p.nbatch, bt);
    NXC_LAUNCH_CHECK(ctx);
    return NXC_OK;
  }
};
template <int DT> struct MmLaunch<DT, false> {
  static nxc_status go(nxc_ctx *, const NxcMatmulProblem &, dim3, const MmBatch &) { return NXC_ERR_UNSUPPORTED_DTYPE; }
};

nxc_status nxc_matmul_simt(nxc_ctx *ctx, const NxcMatmulProblem &p) {
  MmBatch bt;
  bt.nd = p.batch_nd;
  for (int i = 0; i < p.batch_nd; i++) {
    bt.div[i] = nxc_fastdiv_make((uint32_t)p.bshape[i]);
    bt.as_[i] = p.as_[i]; bt.bs_[i] = p.bs_[i]; bt.cs_[i] = p.cs_[i];
  }
  if (p.nbatch >= 0x7FFFFFFFLL) return NXC_ERR_SHAPE;
  // the 128 x 128 kernel once its grid covers the SMs; below that the 64 x 64 tiles of the generic kernel put four
  // times as many CTAs on the device (f32 256^3: 4 CTAs took 61 us)
  const int64_t big_ctas = ((p.m + MB_M - 1) / MB_M) * ((p.n + MB_N - 1) / MB_N) * p.nbatch;
  if ((p.dt == NXC_F32 || p.dt == NXC_F64) && p.m >= 64 && p.n >= 64 && p.k >= 8 && big_ctas >= ctx->sm_count)
    return p.dt == NXC_F32 ? launch_big<float>(ctx, p, bt) : launch_big<double>(ctx, p, bt);
  dim3 grid((unsigned)((p.n + MM_BN - 1) / MM_BN), (unsigned)((p.m + MM_BM - 1) / MM_BM),
            (unsigned)(p.nbatch < 65535 ? p.nbatch : 65535));
  if (grid.y > 65535) return NXC_ERR_SHAPE;
  nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
  NXC_DISPATCH_DTYPE(p.dt, { st = MmLaunch<DT, DT != NXC_BOOL>::go(ctx, p, grid, bt); })
  return st;
}
