// nxc_matmul.cu -- C-ABI matmul entry: validate, resolve batch broadcast, route.
// Replaces caml_nx_c_matmul and its driver's checks (reference:
// nx_c_matmul.c:874-936, 1229-1277): same shape rules, same statuses
// (shape/aliased -> Invalid_argument; dtype mismatch / unsupported -> Failure).
#include "nxc_matmul.cuh"

// The tensor-core kernel feeds on TMA, which wants a unit stride on one matrix dim, 16-byte multiples
// on every other stride and a 16-byte aligned base. A view that cannot be described that way (a row
// pitch of 50257 bf16 -- GPT-2's vocabulary -- a sliced or doubly strided operand, a batch that does
// not collapse) is PACKED: one pass through the backend's own strided copy into a K-major buffer
// with an aligned pitch, then the same kernel. One extra read + write of the operand (HBM-bound)
// against a CUDA-core GEMM 30-40x slower than the tensor cores.
static bool nxc_mm_operand_ok(const char *base, int64_t rs, int64_t cs, int64_t rows, int64_t cols, int esize) {
  const int64_t al = 16 / esize;
  if (((uintptr_t)base & 15) != 0) return false;
  if (cs == 1 || cols == 1) return rs > 0 && rs % al == 0;
  if (rs == 1 || rows == 1) return cs > 0 && cs % al == 0;
  return false;
}

static nxc_status nxc_matmul_tc_packed(nxc_ctx *ctx, const NxcMatmulProblem &q, bool force_a = false) {
  const int esize = (q.dt == NXC_F32) ? 4 : 2;
  const int64_t al = 16 / esize;
  if (q.k == 0 || q.c_cs != 1 || q.m * q.n * q.k < ((int64_t)1 << 24)) return NXC_MM_TC_DECLINED;
  bool a_b = false, b_b = false, bad_batch = false;
  for (int i = 0; i < q.batch_nd; i++) {
    if (q.bshape[i] > 1 && q.as_[i] != 0) a_b = true;
    if (q.bshape[i] > 1 && q.bs_[i] != 0) b_b = true;
    if (q.bshape[i] > 1 && ((q.as_[i] % al) != 0 || (q.bs_[i] % al) != 0)) bad_batch = true;
  }
  // (a batch that does not collapse to one stride is also a reason to pack; detecting it here would
  // repeat the kernel's own walk, so any decline of a large product packs whatever is batched)
  const bool pack_a = force_a || bad_batch || a_b || !nxc_mm_operand_ok(q.a, q.a_rs, q.a_cs, q.m, q.k, esize);
  const bool pack_b = !force_a && (bad_batch || b_b || !nxc_mm_operand_ok(q.b, q.b_cs, q.b_rs, q.n, q.k, esize));
  if (!pack_a && !pack_b) return NXC_MM_TC_DECLINED;  // declined for another reason: nothing to gain
  const int64_t kp = (q.k + al - 1) / al * al;
  NxcMatmulProblem p = q;
  void *pa = NULL, *pb = NULL;
  nxc_status s = NXC_OK;
  auto pack = [&](const char *src, int64_t rows, int64_t rs, int64_t ks, const int64_t *bstr, bool batched, void **out,
                  int64_t *new_bstr) -> nxc_status {
    const int64_t nb = batched ? q.nbatch : 1;
    nxc_status st = nxc_alloc(ctx, (size_t)(nb * rows * kp) * (size_t)esize, out);
    if (st) return st;
    nxc_tensor d, x;
    d.data = *out; x.data = (void *)src;
    d.dtype = x.dtype = q.dt;
    d.offset = x.offset = 0;
    int nd = 0;
    int64_t ext = 1;
    if (batched) {
      for (int i = 0; i < q.batch_nd; i++) { d.shape[nd] = x.shape[nd] = q.bshape[i]; x.strides[nd] = bstr[i]; nd++; }
      for (int i = q.batch_nd - 1; i >= 0; i--) { d.strides[i] = ext * rows * kp; new_bstr[i] = q.bshape[i] > 1 ? d.strides[i] : 0; ext *= q.bshape[i]; }
    } else {
      for (int i = 0; i < q.batch_nd; i++) new_bstr[i] = 0;
    }
    d.shape[nd] = x.shape[nd] = rows; d.strides[nd] = kp; x.strides[nd] = rs; nd++;
    d.shape[nd] = x.shape[nd] = q.k; d.strides[nd] = 1; x.strides[nd] = ks; nd++;
    d.ndim = x.ndim = nd;
    return nxc_copy(ctx, &d, &x);
  };
  if (pack_a) {
    s = pack(q.a, q.m, q.a_rs, q.a_cs, q.as_, a_b, &pa, p.as_);
    p.a = (const char *)pa; p.a_rs = kp; p.a_cs = 1;
  }
  if (!s && pack_b) {
    s = pack(q.b, q.n, q.b_cs, q.b_rs, q.bs_, b_b, &pb, p.bs_);   // B^T, K-major
    p.b = (const char *)pb; p.b_rs = 1; p.b_cs = kp;
  }
  if (!s) s = nxc_matmul_tc(ctx, p);
  if (pa) nxc_free(ctx, pa);   // stream-ordered: reused only after the GEMM that reads them
  if (pb) nxc_free(ctx, pb);
  return s;
}

extern "C" nxc_status nxc_matmul(nxc_ctx *ctx, const nxc_tensor *C, const nxc_tensor *A,
                                 const nxc_tensor *B) {
  NXC_TRACE(ctx, "nxc_matmul");
  nxc_status s = NXC_OK;
  NxcMatmulProblem p;
  if ((s = nxc_check_tensor(A)) || (s = nxc_check_tensor(B)) || (s = nxc_check_tensor(C))) goto fail;
  if (A->dtype != B->dtype || A->dtype != C->dtype) { s = NXC_ERR_DTYPE_MISMATCH; goto fail; }
  {
    const int dt = A->dtype;
    const int cls = nxc_dtype_class(dt);
    if (cls & NXC_CLS_PACKED) { s = NXC_ERR_PACKED; goto fail; }
    if (cls & NXC_CLS_BOOL) { s = NXC_ERR_UNSUPPORTED_DTYPE; goto fail; }
    if (A->ndim < 2 || B->ndim < 2) { s = NXC_ERR_SHAPE; goto fail; }
    const int nd = A->ndim > B->ndim ? A->ndim : B->ndim;
    if (C->ndim != nd) { s = NXC_ERR_SHAPE; goto fail; }
    p.dt = dt;
    p.m = A->shape[A->ndim - 2];
    p.k = A->shape[A->ndim - 1];
    p.n = B->shape[B->ndim - 1];
    if (p.k != B->shape[B->ndim - 2]) { s = NXC_ERR_SHAPE; goto fail; }
    if (C->shape[nd - 2] != p.m || C->shape[nd - 1] != p.n) { s = NXC_ERR_SHAPE; goto fail; }
    // Batch dims broadcast numpy-style: operands are right-aligned against the output's rank, a
    // missing or size-1 dim repeats (stride 0), anything else must agree; the output has the
    // broadcast extent and may not alias itself (rules: nx_c_matmul.c:895-926).
    p.batch_nd = nd - 2;
    p.nbatch = 1;
    struct BatchDim { int64_t extent, stride; };
    auto batch_dim = [nd](const nxc_tensor *t, int i) -> BatchDim {
      const int j = i - (nd - t->ndim);  // this operand's own dim index, negative = absent
      if (j < 0 || t->shape[j] == 1) return {1, 0};
      return {t->shape[j], t->strides[j]};
    };
    for (int i = 0; i < p.batch_nd; i++) {
      const BatchDim da = batch_dim(A, i), db = batch_dim(B, i);
      const int64_t extent = da.extent > db.extent ? da.extent : db.extent;
      const bool agree = da.extent == db.extent || da.extent == 1 || db.extent == 1;
      if (!agree || C->shape[i] != extent) { s = NXC_ERR_SHAPE; goto fail; }
      if (extent > 1 && C->strides[i] == 0) { s = NXC_ERR_OUT_ALIASED; goto fail; }
      p.bshape[i] = extent;
      p.as_[i] = da.stride;
      p.bs_[i] = db.stride;
      p.cs_[i] = C->strides[i];
      p.nbatch *= extent;
    }
    if (p.m == 0 || p.n == 0 || p.nbatch == 0) return NXC_OK;
    p.a_rs = A->strides[A->ndim - 2]; p.a_cs = A->strides[A->ndim - 1];
    p.b_rs = B->strides[B->ndim - 2]; p.b_cs = B->strides[B->ndim - 1];
    p.c_rs = C->strides[nd - 2]; p.c_cs = C->strides[nd - 1];
    if ((p.m > 1 && p.c_rs == 0) || (p.n > 1 && p.c_cs == 0)) { s = NXC_ERR_OUT_ALIASED; goto fail; }
    const int64_t es = nxc_elem_size(dt);
    p.a = (const char *)A->data + A->offset * es;
    p.b = (const char *)B->data + B->offset * es;
    p.c = (char *)C->data + C->offset * es;

    // x [batch..., m, k] @ W [k, n] -- every Kaun Linear.apply on a [B, T, C] activation (linear.ml:50-52)
    // -- is ONE product with batch * m rows when the batches of x and of the output follow each
    // other at the row pitch: fold them, so the kernels see a tall matrix (full tiles, one
    // launch-worth of work) instead of `batch` short ones (the GPT-2 step at 4 x 64 tokens ran its
    // linears as 4 products of 64 rows: 24 CTAs on 148 SMs).
    if (p.nbatch > 1 && p.m > 0) {
      bool fold = true;
      int64_t rows = p.m;
      for (int i = p.batch_nd - 1; i >= 0 && fold; i--) {
        if (p.bshape[i] == 1) continue;
        fold = p.bs_[i] == 0 && p.as_[i] == rows * p.a_rs && p.cs_[i] == rows * p.c_rs;
        rows *= p.bshape[i];
      }
      if (fold) { p.m = rows; p.nbatch = 1; p.batch_nd = 0; }
    }

    const bool tc_dtype = (dt == NXC_BF16 || dt == NXC_F16 || (dt == NXC_F32 && ctx->matmul_tf32 == 1));
    if (tc_dtype) {
      // An M-major LEFT operand (a transposed view: the x^T of every dW = x^T g) is consumed in place,
      // but the tensor pipe then runs at 59 % instead of 77 % active (ncu, 8192^3: 951 vs 800 us with
      // identical DRAM / L2 traffic; an N-major right operand costs nothing). With N >= 4096 one
      // transposing pass over A (HBM-bound, 2 M K elements moved) is cheaper than that.
      s = NXC_MM_TC_DECLINED;
      if (p.a_rs == 1 && p.a_cs != 1 && p.m > 1 && p.nbatch == 1 && p.n >= 4096 && p.m * p.k >= ((int64_t)1 << 22) &&
          !getenv("NX_CUDA_MM_NO_APACK"))
        s = nxc_matmul_tc_packed(ctx, p, /*force_a=*/true);
      if (s == NXC_MM_TC_DECLINED) s = nxc_matmul_tc(ctx, p);
      if (s == NXC_MM_TC_DECLINED) s = nxc_matmul_tc_packed(ctx, p);
      if (s != NXC_MM_TC_DECLINED) { if (s) goto fail; return NXC_OK; }
    }
    // f32 at f32-class accuracy on the tensor cores (3xTF32, nxc_matmul_x3.cu) -- the default for
    // f32 once the product is a quarter GFLOP and so worth the two split passes (a 256 x 768 x 768
    // linear of the GPT-2 step: 260 us on the CUDA-core kernel, whose 128 x 128 tiles leave most SMs
    // idle at that size; measured 8192^3: 230
    // vs 30 TFLOP/s; error 2e-6 .. 6e-5 of max |A||B| for K = 1024 .. 8192, inside the classical
    // K*u sgemm bound and 20x inside the reference's own f32 matmul tolerance, 1e-3 rel + 1e-3 abs,
    // backend_c/test/matmul_test.ml:831). Mode "ieee" keeps every f32 product on the CUDA-core
    // kernel (each product and sum rounded to nearest, like the reference's microkernel).
    if (dt == NXC_F32 && (ctx->matmul_tf32 == 0 || ctx->matmul_tf32 == 2) && p.m >= 32 && p.n >= 32 &&
        2.0 * (double)p.m * (double)p.n * (double)p.k * (double)p.nbatch >= 268435456.0) {
      s = nxc_matmul_f32x3(ctx, p);
      if (s != NXC_MM_TC_DECLINED) { if (s) goto fail; return NXC_OK; }
    }
    s = nxc_matmul_simt(ctx, p);
    if (s) goto fail;
    return NXC_OK;
  }
fail:
  if (s && strcmp(s, NXC_ERR_CUDA) != 0) snprintf(ctx->err, sizeof ctx->err, "%s", s);
  return s;
}
