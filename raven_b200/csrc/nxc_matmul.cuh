// nxc_matmul.cuh -- the resolved matmul problem shared by the SIMT and tcgen05 paths.
#pragma once
#include "nxc_common.cuh"

struct NxcMatmulProblem {
  int dt;
  int64_t m, n, k;
  int64_t a_rs, a_cs, b_rs, b_cs, c_rs, c_cs;  // element strides
  int batch_nd;
  int64_t nbatch;
  int64_t bshape[NXC_MAX_NDIM], as_[NXC_MAX_NDIM], bs_[NXC_MAX_NDIM], cs_[NXC_MAX_NDIM];
  const char *a;  // first live element of each operand
  const char *b;
  char *c;
};

// Generic CUDA-core GEMM for every compute dtype, arbitrary strides.
nxc_status nxc_matmul_simt(nxc_ctx *ctx, const NxcMatmulProblem &p);
// tcgen05/TMEM/TMA GEMM for bf16 / f16 (and f32 in tf32 mode). Returns
// NXC_MM_TC_DECLINED (not an error) when the layout cannot be described to TMA;
// the caller then packs the operands and retries, or uses the SIMT path.
#define NXC_MM_TC_DECLINED ((nxc_status) "tc:declined")
nxc_status nxc_matmul_tc(nxc_ctx *ctx, const NxcMatmulProblem &p);
// f32 operands split into tf32 hi / lo parts and multiplied as one tf32 GEMM over a tripled K
// axis (nxc_matmul_x3.cu): f32-class accuracy at tensor-core speed. May decline like the above.
nxc_status nxc_matmul_f32x3(nxc_ctx *ctx, const NxcMatmulProblem &p);
