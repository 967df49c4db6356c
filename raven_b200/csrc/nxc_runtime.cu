// nxc_runtime.cu -- context, stream, stream-ordered caching allocator,
// transfers, status classification and the host-side map plan.
//
// Replaces, for a device: `create_context` (reference:
// backend/nx_backend.mli:32-39), Nx_buffer.create / to_host / from_host as used
// by the veneer (reference: backend_c/nx_backend.ml:50-69), the status ->
// exception classifier (reference: nx_c_engine.c:1345-1351) and the dimension
// coalescer (reference: nx_c_engine.c:542-590). The reference's pthread pool
// has no equivalent here: the grid is the pool.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nxc_map.cuh"

nxc_status nxc_cuda_fail(nxc_ctx *ctx, cudaError_t e, const char *what) {
  if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s: %s (%s)", NXC_ERR_CUDA, cudaGetErrorString(e), what);
  cudaGetLastError();  // clear sticky-less errors so the next call starts clean
  if (e == cudaErrorMemoryAllocation) return NXC_ERR_ALLOC;
  return NXC_ERR_CUDA;
}

static char g_create_err[512] = "";

extern "C" nxc_status nxc_ctx_create_on(int device, void *cuda_stream, nxc_ctx **out) {
  *out = NULL;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    snprintf(g_create_err, sizeof g_create_err, "%s: %s", NXC_ERR_NO_DEVICE,
             e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
    cudaGetLastError();
    return NXC_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) device = 0;
  nxc_ctx *ctx = (nxc_ctx *)calloc(1, sizeof(nxc_ctx));
  if (!ctx) return NXC_ERR_ALLOC;
  ctx->device = device;
  ctx->err[0] = 0;
  NXC_CUDA_TRY(ctx, cudaSetDevice(device));
  if (cuda_stream) {
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
  } else {
    NXC_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  cudaDeviceProp prop;
  NXC_CUDA_TRY(ctx, cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  // Stream-ordered allocator as the caching allocator: never trim the pool, so a
  // free()d block is reused by the next alloc on the stream without touching the
  // driver -- every eager op allocates a fresh output (reference:
  // backend_c/nx_backend.ml:54-57), so this is on the per-op critical path.
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  const char *mm = getenv("NX_CUDA_MATMUL");
  ctx->matmul_tf32 = (mm && strcmp(mm, "tf32") == 0) ? 1 : (mm && strcmp(mm, "ieee") == 0) ? 3 : 0;
  ctx->rank = 0;
  ctx->world = 1;
  // TMA descriptor encoder: resolved through the runtime so the library has no
  // link-time dependency on libcuda (it must dlopen on a GPU-less build box).
  {
    void *fn = NULL;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      ctx->encode_tiled = fn;
    cudaGetLastError();
  }
  *out = ctx;
  return NXC_OK;
}

extern "C" nxc_status nxc_ctx_create(nxc_ctx **out) {
  int dev = 0;
  const char *s = getenv("NX_CUDA_DEVICE");
  if (!s) s = getenv("LOCAL_RANK");
  if (s) dev = atoi(s);
  return nxc_ctx_create_on(dev, NULL, out);
}

// ---- side streams and deferred frees ----------------------------------------------------------
nxc_status nxc_side_streams(nxc_ctx *ctx) {
  if (ctx->h2d_stream) return NXC_OK;
  NXC_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
  NXC_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  NXC_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  NXC_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  NXC_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_h2d, cudaEventDisableTiming));
  NXC_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming));
  return NXC_OK;
}
// Release the buffers whose read-back has finished (`drain`: wait for the rest first).
static void nxc_reap_pending(nxc_ctx *ctx, bool drain) {
  int w = 0;
  for (int i = 0; i < ctx->n_pending; i++) {
    nxc_ctx::nxc_pending &e = ctx->pending[i];
    bool done = drain ? (cudaEventSynchronize(e.done) == cudaSuccess, true) : (cudaEventQuery(e.done) == cudaSuccess);
    if (done) {
      if (e.freed) cudaFreeAsync(e.ptr, ctx->stream);
      cudaEventDestroy(e.done);
    } else {
      ctx->pending[w++] = e;
    }
  }
  ctx->n_pending = w;
  cudaGetLastError();  // cudaErrorNotReady from the queries is not an error
}

extern "C" void nxc_ctx_destroy(nxc_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->h2d_stream) {
    cudaStreamSynchronize(ctx->h2d_stream);
    cudaStreamSynchronize(ctx->d2h_stream);
    cudaStreamSynchronize(ctx->comm_stream);
    nxc_reap_pending(ctx, true);
    cudaStreamDestroy(ctx->h2d_stream);
    cudaStreamDestroy(ctx->d2h_stream);
    cudaStreamDestroy(ctx->comm_stream);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_h2d);
    cudaEventDestroy(ctx->ev_comm);
  }
  free(ctx->pending);
  if (ctx->scratch) cudaFreeAsync(ctx->scratch, ctx->stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  free(ctx);
}

extern "C" nxc_status nxc_sync(nxc_ctx *ctx) {
  NXC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->h2d_stream) {
    NXC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->h2d_stream));
    NXC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->d2h_stream));
    NXC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->comm_stream));
    nxc_reap_pending(ctx, true);
  }
  return NXC_OK;
}
extern "C" void *nxc_stream(nxc_ctx *ctx) { return (void *)ctx->stream; }
extern "C" int nxc_device(nxc_ctx *ctx) { return ctx->device; }
extern "C" const char *nxc_last_error(nxc_ctx *ctx) { return ctx ? ctx->err : g_create_err; }
extern "C" uint64_t nxc_launch_count(nxc_ctx *ctx) { return ctx->launches; }

extern "C" int nxc_set_matmul_mode(nxc_ctx *ctx, const char *mode) {
  if (strcmp(mode, "tf32") == 0) { ctx->matmul_tf32 = 1; return 0; }
  if (strcmp(mode, "f32") == 0) { ctx->matmul_tf32 = 0; return 0; }
  if (strcmp(mode, "f32x3") == 0) { ctx->matmul_tf32 = 2; return 0; }
  if (strcmp(mode, "ieee") == 0) { ctx->matmul_tf32 = 3; return 0; }
  return -1;
}

extern "C" int nxc_status_is_invalid_argument(nxc_status s) {
  if (!s) return 0;
  static const char *inv[] = {NXC_ERR_EMPTY_REDUCE, NXC_ERR_AXES, NXC_ERR_AXIS,
                              NXC_ERR_OUT_RANK, NXC_ERR_OUT_ALIASED, NXC_ERR_SHAPE,
                              "threefry: last axis must have extent 2",
                              // linalg preconditions (reference: la_raise, nx_c_linalg.h:310-318)
                              "linalg requires a float or complex dtype", "matrix must be square",
                              "eig requires a float or complex dtype",
                              "operand shapes are incompatible"};
  for (size_t i = 0; i < sizeof inv / sizeof inv[0]; i++)
    if (strcmp(s, inv[i]) == 0) return 1;
  return 0;
}

extern "C" int64_t nxc_elem_size(int dt) {
  switch (dt) {
    case NXC_F8E4M3: case NXC_F8E5M2: case NXC_I8: case NXC_U8: case NXC_BOOL: return 1;
    case NXC_F16: case NXC_BF16: case NXC_I16: case NXC_U16: return 2;
    case NXC_F32: case NXC_I32: case NXC_U32: return 4;
    case NXC_F64: case NXC_I64: case NXC_U64: case NXC_C32: return 8;
    case NXC_C64: return 16;
    default: return 0;
  }
}

extern "C" nxc_status nxc_alloc(nxc_ctx *ctx, size_t bytes, void **dptr) {
  *dptr = NULL;
  if (bytes == 0) bytes = 16;
  if (ctx->n_pending) nxc_reap_pending(ctx, false);
  NXC_CUDA_TRY(ctx, cudaMallocAsync(dptr, bytes, ctx->stream));
  return NXC_OK;
}
extern "C" nxc_status nxc_free(nxc_ctx *ctx, void *dptr) {
  if (!dptr) return NXC_OK;
  if (ctx->n_pending) {
    nxc_reap_pending(ctx, false);
    for (int i = 0; i < ctx->n_pending; i++)
      if (ctx->pending[i].ptr == dptr) {  // a read-back still owns it: released by nxc_reap_pending
        ctx->pending[i].freed = 1;
        return NXC_OK;
      }
  }
  NXC_CUDA_TRY(ctx, cudaFreeAsync(dptr, ctx->stream));
  return NXC_OK;
}
extern "C" nxc_status nxc_host_alloc(nxc_ctx *ctx, size_t bytes, void **hptr) {
  *hptr = NULL;
  NXC_CUDA_TRY(ctx, cudaMallocHost(hptr, bytes ? bytes : 16));
  return NXC_OK;
}
extern "C" nxc_status nxc_host_free(nxc_ctx *ctx, void *hptr) {
  if (hptr) NXC_CUDA_TRY(ctx, cudaFreeHost(hptr));
  return NXC_OK;
}
static bool nxc_is_pinned(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}
extern "C" nxc_status nxc_h2d(nxc_ctx *ctx, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return NXC_OK;
  static const bool engines = !(getenv("NX_CUDA_COPY_ENGINES") && getenv("NX_CUDA_COPY_ENGINES")[0] == '0');
  if (engines && bytes >= ((size_t)1 << 20) && nxc_is_pinned(src)) {
    // upload engine: after what is queued (dst's allocation included), before what follows
    nxc_status s = nxc_side_streams(ctx);
    if (s) return s;
    NXC_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    NXC_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->h2d_stream, ctx->ev_fork, 0));
    NXC_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->h2d_stream));
    NXC_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_h2d, ctx->h2d_stream));
    NXC_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d, 0));
    return NXC_OK;
  }
  NXC_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return NXC_OK;
}
extern "C" nxc_status nxc_d2h_async(nxc_ctx *ctx, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return NXC_OK;
  if (!nxc_is_pinned(dst)) {
    snprintf(ctx->err, sizeof ctx->err, "%s: nxc_d2h_async needs pinned host memory (nxc_host_alloc)", NXC_ERR_CUDA);
    return NXC_ERR_CUDA;
  }
  nxc_status s = nxc_side_streams(ctx);
  if (s) return s;
  NXC_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
  NXC_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_fork, 0));
  NXC_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
  if (ctx->n_pending == ctx->cap_pending) {
    int cap = ctx->cap_pending ? 2 * ctx->cap_pending : 8;
    void *np = realloc(ctx->pending, cap * sizeof *ctx->pending);
    if (!np) return NXC_ERR_ALLOC;
    ctx->pending = (nxc_ctx::nxc_pending *)np;
    ctx->cap_pending = cap;
  }
  nxc_ctx::nxc_pending &e = ctx->pending[ctx->n_pending];
  e.ptr = (void *)src;
  e.freed = 0;
  NXC_CUDA_TRY(ctx, cudaEventCreateWithFlags(&e.done, cudaEventDisableTiming));
  NXC_CUDA_TRY(ctx, cudaEventRecord(e.done, ctx->d2h_stream));
  ctx->n_pending++;
  return NXC_OK;
}
extern "C" nxc_status nxc_d2h(nxc_ctx *ctx, void *dst, const void *src, size_t bytes) {
  if (bytes) NXC_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  NXC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return NXC_OK;
}
extern "C" nxc_status nxc_memset(nxc_ctx *ctx, void *dst, int byte, size_t bytes) {
  if (bytes) NXC_CUDA_TRY(ctx, cudaMemsetAsync(dst, byte, bytes, ctx->stream));
  return NXC_OK;
}

nxc_status nxc_scratch(nxc_ctx *ctx, size_t bytes, void **out) {
  if (bytes > ctx->scratch_bytes) {
    if (ctx->scratch) NXC_CUDA_TRY(ctx, cudaFreeAsync(ctx->scratch, ctx->stream));
    ctx->scratch = NULL;
    ctx->scratch_bytes = 0;
    size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
    NXC_CUDA_TRY(ctx, cudaMallocAsync(&ctx->scratch, want, ctx->stream));
    ctx->scratch_bytes = want;
  }
  *out = ctx->scratch;
  return NXC_OK;
}

// ---- the map plan ---------------------------------------------------------------
nxc_status nxc_map_plan(const nxc_tensor *const *ops, int nop, const int64_t *esize,
                        NxcMapPlan *p) {
  if (nop > NXC_MAX_OPERANDS) return NXC_ERR_ARITY;
  const int ndim = ops[0]->ndim;
  p->nop = nop;
  for (int k = 0; k < nop; k++) {
    if (ops[k]->ndim != ndim) return NXC_ERR_RANK_MISMATCH;
    p->base[k] = (char *)ops[k]->data + ops[k]->offset * esize[k];
  }
  int64_t total = 1;
  for (int i = 0; i < ndim; i++) total *= ops[0]->shape[i];
  p->total = total;
  int nd = 0;
  for (int i = 0; i < ndim; i++) {
    const int64_t s = ops[0]->shape[i];
    if (s == 1) continue;
    bool merge = nd > 0;
    for (int k = 0; k < nop && merge; k++)
      if (p->stride[k][nd - 1] != ops[k]->strides[i] * s) merge = false;
    if (merge) {
      p->shape[nd - 1] *= s;
      for (int k = 0; k < nop; k++) p->stride[k][nd - 1] = ops[k]->strides[i];
    } else {
      p->shape[nd] = s;
      for (int k = 0; k < nop; k++) p->stride[k][nd] = ops[k]->strides[i];
      nd++;
    }
  }
  if (nd == 0) {
    p->shape[0] = 1;
    for (int k = 0; k < nop; k++) p->stride[k][0] = (k == 0) ? 1 : 0;
    nd = 1;
  }
  p->ndim = nd;
  if (total != 0)
    for (int i = 0; i < nd; i++)
      if (p->shape[i] > 1 && p->stride[0][i] == 0) return NXC_ERR_OUT_ALIASED;
  return NXC_OK;
}
