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
  dim3 grid((unsigned)((p.n + MM_BN - 1) / MM_BN), (unsigned)((p.m + MM_BM - 1) / MM_BM),
            (unsigned)(p.nbatch < 65535 ? p.nbatch : 65535));
  if (grid.y > 65535) return NXC_ERR_SHAPE;
  nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
  NXC_DISPATCH_DTYPE(p.dt, { st = MmLaunch<DT, DT != NXC_BOOL>::go(ctx, p, grid, bt); })
  return st;
}
