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

#include <map>
#include <vector>

#include "nxc_map.cuh"

// ---- step capture: the graph and its arena -----------------------------------------------------
// An eager backend issues one launch per op from the host; a training step is a few thousand of
// them and a sharded reduction step couples the ranks' hosts through the exchange kernels. A
// captured step is replayed by ONE cudaGraphLaunch. What makes the eager op sequence capturable
// is memory: every op allocates its output (reference: backend_c/nx_backend.ml:54-57), and the
// stream-ordered pool would turn those into graph memory nodes whose lifetime rules do not fit
// handles the host keeps after the capture. So while capturing, nxc_alloc / nxc_free are served by
// an arena the graph owns: plain cudaMalloc chunks, a first-fit free list with coalescing. The
// capture is one linear chain on the context stream (side-stream work is joined back at once), so
// reusing a block freed earlier in the capture is ordered exactly as it was in the eager run. The
// handles created during the capture stay valid -- they are the replay's outputs -- until the
// graph is destroyed; nxc_free of them outside the capture is a no-op.
struct nxc_graph {
  cudaGraph_t graph = NULL;
  cudaGraphExec_t exec = NULL;
  struct Chunk { char *base; size_t size; };
  struct Free { size_t size; int chunk; };
  std::vector<Chunk> chunks;
  std::map<char *, Free> free_;                     // by address: neighbours coalesce
  std::map<char *, std::pair<size_t, int>> live;    // ptr -> (size, chunk)
  uint64_t kernels = 0;                             // kernel nodes = launches per replay
  size_t arena_bytes = 0, peak_live = 0, now_live = 0;
  size_t internal = 0;                              // live blocks the engine itself holds (scratch): not handles
  bool destroyed = false;                           // nxc_graph_destroy was called; the arena waits for its handles
  bool owns(const void *p) const {
    for (const Chunk &c : chunks)
      if ((const char *)p >= c.base && (const char *)p < c.base + c.size) return true;
    return false;
  }
};

static nxc_status arena_alloc(nxc_ctx *ctx, nxc_graph *g, size_t bytes, void **out) {
  const size_t need = (bytes + 511) & ~(size_t)511;
  for (auto it = g->free_.begin(); it != g->free_.end(); ++it) {
    if (it->second.size < need) continue;
    char *p = it->first;
    const nxc_graph::Free f = it->second;
    g->free_.erase(it);
    if (f.size > need) g->free_[p + need] = {f.size - need, f.chunk};
    g->live[p] = {need, f.chunk};
    g->now_live += need;
    if (g->now_live > g->peak_live) g->peak_live = g->now_live;
    *out = p;
    return NXC_OK;
  }
  // a new chunk: cudaMalloc is legal here because the capture runs in relaxed mode
  size_t want = need < ((size_t)64 << 20) ? ((size_t)64 << 20) : need;
  void *base = NULL;
  cudaError_t e = cudaMalloc(&base, want);
  if (e != cudaSuccess && want > need) { cudaGetLastError(); want = need; e = cudaMalloc(&base, want); }
  if (e != cudaSuccess) return nxc_cuda_fail(ctx, e, "cudaMalloc (capture arena)");
  const int ci = (int)g->chunks.size();
  g->chunks.push_back({(char *)base, want});
  g->arena_bytes += want;
  if (want > need) g->free_[(char *)base + need] = {want - need, ci};
  g->live[(char *)base] = {need, ci};
  g->now_live += need;
  if (g->now_live > g->peak_live) g->peak_live = g->now_live;
  *out = base;
  return NXC_OK;
}
static void arena_free(nxc_graph *g, void *ptr) {
  auto it = g->live.find((char *)ptr);
  if (it == g->live.end()) return;
  char *p = it->first;
  size_t size = it->second.first;
  const int chunk = it->second.second;
  g->live.erase(it);
  g->now_live -= size;
  auto nx = g->free_.lower_bound(p);
  if (nx != g->free_.end() && nx->second.chunk == chunk && p + size == nx->first) {
    size += nx->second.size;
    nx = g->free_.erase(nx);
  }
  if (nx != g->free_.begin()) {
    auto pv = std::prev(nx);
    if (pv->second.chunk == chunk && pv->first + pv->second.size == p) {
      pv->second.size += size;
      return;
    }
  }
  g->free_[p] = {size, chunk};
}
static void graph_forget(nxc_ctx *ctx, nxc_graph *g);
static nxc_graph *graph_owning(nxc_ctx *ctx, const void *p) {
  for (int i = 0; i < ctx->n_graphs; i++)
    if (ctx->graphs[i]->owns(p)) return ctx->graphs[i];
  return NULL;
}

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
  {
    void *hp = NULL, *dp = NULL;
    NXC_CUDA_TRY(ctx, cudaHostAlloc(&hp, NXC_ST_WORDS * sizeof(int), cudaHostAllocMapped));
    memset(hp, 0, NXC_ST_WORDS * sizeof(int));
    NXC_CUDA_TRY(ctx, cudaHostGetDevicePointer(&dp, hp, 0));
    ctx->hstatus = (volatile int *)hp;
    ctx->dstatus = (int *)dp;
  }
  {
    const char *nv = getenv("NX_CUDA_NVTX");
    ctx->nvtx = (nv && nv[0] == '1') ? 1 : 0;
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
      // several read-backs of one buffer: the copies finish in issue order (one stream), so the
      // buffer goes back to the pool with the LAST entry that names it
      bool later = false;
      for (int j = i + 1; j < ctx->n_pending && !later; j++) later = ctx->pending[j].ptr == e.ptr;
      if (e.freed && !later) cudaFreeAsync(e.ptr, ctx->stream);
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
  while (ctx->n_graphs > 0) graph_forget(ctx, ctx->graphs[ctx->n_graphs - 1]);
  free(ctx->graphs);
  if (ctx->hstatus) cudaFreeHost((void *)ctx->hstatus);
  if (ctx->scratch) cudaFreeAsync(ctx->scratch, ctx->stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  free(ctx);
}

nxc_status nxc_capture_refuse(nxc_ctx *ctx, const char *what) {
  snprintf(ctx->err, sizeof ctx->err, "%s: %s", NXC_ERR_CAPTURE, what);
  return NXC_ERR_CAPTURE;
}
// What the kernels left in the status page since the last report. Called once the stream has
// drained (the writes of finished kernels to mapped host memory are visible then).
nxc_status nxc_status_page_check(nxc_ctx *ctx) {
  if (!ctx->hstatus) return NXC_OK;
  if (ctx->hstatus[NXC_ST_EXCHANGE]) {
    ctx->hstatus[NXC_ST_EXCHANGE] = 0;
    ctx->dist_poisoned = 1;
    snprintf(ctx->err, sizeof ctx->err, "%s: %s (the results of that exchange are undefined; the communicator "
             "is unusable from here on)", NXC_ERR_NCCL, NXC_ERR_EXCHANGE_TIMEOUT);
    return NXC_ERR_NCCL;
  }
  if (ctx->hstatus[NXC_ST_INDEX]) {
    ctx->hstatus[NXC_ST_INDEX] = 0;
    snprintf(ctx->err, sizeof ctx->err, "%s", NXC_ERR_INDEX_OOB);
    return NXC_ERR_INDEX_OOB;
  }
  return NXC_OK;
}
extern "C" nxc_status nxc_sync(nxc_ctx *ctx) {
  if (nxc_is_capturing(ctx)) return nxc_capture_refuse(ctx, "nxc_sync");
  NXC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->h2d_stream) {
    NXC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->h2d_stream));
    NXC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->d2h_stream));
    NXC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->comm_stream));
    nxc_reap_pending(ctx, true);
  }
  return nxc_status_page_check(ctx);
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
  if (nxc_is_capturing(ctx)) return arena_alloc(ctx, ctx->capturing, bytes, dptr);
  if (ctx->n_pending) nxc_reap_pending(ctx, false);
  NXC_CUDA_TRY(ctx, cudaMallocAsync(dptr, bytes, ctx->stream));
  return NXC_OK;
}
extern "C" nxc_status nxc_free(nxc_ctx *ctx, void *dptr) {
  if (!dptr) return NXC_OK;
  if (nxc_is_capturing(ctx) && ctx->capturing->owns(dptr)) {  // reusable by the rest of the capture
    arena_free(ctx->capturing, dptr);
    return NXC_OK;
  }
  if (ctx->n_graphs) {
    // A handle created during a capture (a replay's output). The graph owns the memory; the handle's
    // release is only COUNTED, so that a graph destroyed before its handles (finalisers run in no
    // particular order under a GC) keeps its arena until the last of them is gone -- a stale arena
    // address must never reach cudaFreeAsync, it may by then belong to somebody else.
    if (nxc_graph *g = graph_owning(ctx, dptr)) {
      g->live.erase((char *)dptr);
      if (g->destroyed && g->live.size() <= g->internal) graph_forget(ctx, g);
      return NXC_OK;
    }
  }
  if (ctx->n_pending) {
    nxc_reap_pending(ctx, false);
    bool owned = false;
    for (int i = 0; i < ctx->n_pending; i++)
      if (ctx->pending[i].ptr == dptr) {  // read-backs still own it: released by nxc_reap_pending
        ctx->pending[i].freed = 1;
        owned = true;
      }
    if (owned) return NXC_OK;
  }
  if (nxc_is_capturing(ctx)) {
    // a pool buffer dropped in mid-capture: cudaFreeAsync here would become a graph node that frees
    // it again on every replay. Release it once the capture has ended instead.
    ctx->capturing->live[(char *)dptr] = {0, -1};
    return NXC_OK;
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
  if (nxc_is_capturing(ctx)) {
    // a replay re-reads the host buffer: only page-locked memory may be named by a copy node
    if (!nxc_is_pinned(src)) return nxc_capture_refuse(ctx, "nxc_h2d from pageable host memory");
    NXC_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return NXC_OK;
  }
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
  if (nxc_is_capturing(ctx)) {
    // inside a captured step the copy is a node of the one chain (the arena may hand the source
    // block to a later op of the same capture, so nothing may run beside it)
    NXC_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return NXC_OK;
  }
  nxc_status s = nxc_side_streams(ctx);
  if (s) return s;
  // the tracking entry first: once the copy is queued it must not go untracked
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
  cudaError_t ce = cudaEventRecord(ctx->ev_fork, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_fork, 0);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream);
  if (ce == cudaSuccess) ce = cudaEventRecord(e.done, ctx->d2h_stream);
  if (ce != cudaSuccess) {
    cudaStreamSynchronize(ctx->d2h_stream);  // whatever part was queued no longer reads src after this
    cudaEventDestroy(e.done);
    return nxc_cuda_fail(ctx, ce, "nxc_d2h_async");
  }
  ctx->n_pending++;
  return NXC_OK;
}
extern "C" nxc_status nxc_d2h(nxc_ctx *ctx, void *dst, const void *src, size_t bytes) {
  if (nxc_is_capturing(ctx)) return nxc_capture_refuse(ctx, "nxc_d2h (blocking read-back)");
  if (bytes) NXC_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  NXC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return nxc_status_page_check(ctx);
}
extern "C" nxc_status nxc_memset(nxc_ctx *ctx, void *dst, int byte, size_t bytes) {
  if (bytes) NXC_CUDA_TRY(ctx, cudaMemsetAsync(dst, byte, bytes, ctx->stream));
  return NXC_OK;
}

nxc_status nxc_scratch(nxc_ctx *ctx, size_t bytes, void **out) {
  if (bytes > ctx->scratch_bytes && nxc_is_capturing(ctx)) {
    // growing inside a capture: the larger block comes from the graph's arena and serves the rest
    // of the capture; the context's own scratch is put back by nxc_capture_end
    void *p = NULL;
    size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
    nxc_status s = arena_alloc(ctx, ctx->capturing, want, &p);
    if (s) return s;
    ctx->capturing->internal++;
    ctx->scratch = p;
    ctx->scratch_bytes = want;
  }
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

// ---- step capture --------------------------------------------------------------------------------
extern "C" nxc_status nxc_capture_begin(nxc_ctx *ctx) {
  if (nxc_is_capturing(ctx)) return nxc_capture_refuse(ctx, "nxc_capture_begin (already capturing)");
  nxc_status s = nxc_side_streams(ctx);  // created outside the capture
  if (s) return s;
  if (ctx->n_pending) nxc_reap_pending(ctx, false);
  nxc_graph *g = new nxc_graph();
  ctx->saved_scratch = ctx->scratch;
  ctx->saved_scratch_bytes = ctx->scratch_bytes;
  // relaxed: other threads of the process (NCCL's proxy, torch's watchdog) keep making CUDA calls,
  // and the arena calls cudaMalloc from this one
  cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed);
  if (e != cudaSuccess) { delete g; return nxc_cuda_fail(ctx, e, "cudaStreamBeginCapture"); }
  ctx->capturing = g;
  return NXC_OK;
}

static void graph_release(nxc_ctx *ctx, nxc_graph *g) {
  (void)ctx;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  for (const nxc_graph::Chunk &c : g->chunks) cudaFree(c.base);
  cudaGetLastError();
  delete g;
}
// drop a graph from the context's list and free its arena
static void graph_forget(nxc_ctx *ctx, nxc_graph *g) {
  for (int i = 0; i < ctx->n_graphs; i++)
    if (ctx->graphs[i] == g) {
      ctx->graphs[i] = ctx->graphs[--ctx->n_graphs];
      break;
    }
  graph_release(ctx, g);
}

extern "C" nxc_status nxc_capture_end(nxc_ctx *ctx, nxc_graph **out) {
  *out = NULL;
  nxc_graph *g = ctx->capturing;
  if (!g) return nxc_capture_refuse(ctx, "nxc_capture_end without nxc_capture_begin");
  // async collectives issued during the capture forked the communication stream: join it
  cudaError_t e = cudaSuccess;
  if (ctx->comm_stream) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->comm_stream, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive) {
      e = cudaEventRecord(ctx->ev_comm, ctx->comm_stream);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0);
    }
    cudaGetLastError();
  }
  cudaError_t ee = cudaStreamEndCapture(ctx->stream, &g->graph);
  ctx->capturing = NULL;
  ctx->scratch = ctx->saved_scratch;
  ctx->scratch_bytes = ctx->saved_scratch_bytes;
  // pool buffers dropped during the capture (recorded with chunk -1) go back to the pool now
  for (auto it = g->live.begin(); it != g->live.end();) {
    if (it->second.second < 0) { cudaFreeAsync(it->first, ctx->stream); it = g->live.erase(it); }
    else ++it;
  }
  if (e == cudaSuccess) e = ee;
  if (e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) {
    graph_release(ctx, g);
    return nxc_cuda_fail(ctx, e, "nxc_capture_end");
  }
  size_t n = 0;
  if (cudaGraphGetNodes(g->graph, NULL, &n) == cudaSuccess && n > 0) {
    std::vector<cudaGraphNode_t> nodes(n);
    if (cudaGraphGetNodes(g->graph, nodes.data(), &n) == cudaSuccess)
      for (size_t i = 0; i < n; i++) {
        cudaGraphNodeType t;
        if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) g->kernels++;
      }
  }
  cudaGetLastError();
  if (ctx->n_graphs == ctx->cap_graphs) {
    int cap = ctx->cap_graphs ? 2 * ctx->cap_graphs : 4;
    void *np = realloc(ctx->graphs, cap * sizeof *ctx->graphs);
    if (!np) { graph_release(ctx, g); return NXC_ERR_ALLOC; }
    ctx->graphs = (nxc_graph **)np;
    ctx->cap_graphs = cap;
  }
  ctx->graphs[ctx->n_graphs++] = g;
  *out = g;
  return NXC_OK;
}

extern "C" nxc_status nxc_graph_launch(nxc_ctx *ctx, nxc_graph *g) {
  if (nxc_is_capturing(ctx)) return nxc_capture_refuse(ctx, "nxc_graph_launch");
  if (!g || g->destroyed || !g->exec) {
    snprintf(ctx->err, sizeof ctx->err, "%s: nxc_graph_launch of a destroyed graph", NXC_ERR_CUDA);
    return NXC_ERR_CUDA;
  }
  NXC_CUDA_TRY(ctx, cudaGraphLaunch(g->exec, ctx->stream));
  ctx->launches += g->kernels;
  return NXC_OK;
}
extern "C" uint64_t nxc_graph_kernels(nxc_graph *g) { return g->kernels; }
extern "C" size_t nxc_graph_arena_bytes(nxc_graph *g) { return g->arena_bytes; }

extern "C" void nxc_graph_destroy(nxc_ctx *ctx, nxc_graph *g) {
  if (!g || g->destroyed) return;
  cudaStreamSynchronize(ctx->stream);  // no replay may still be running on the arena
  g->destroyed = true;
  if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = NULL; }
  if (g->graph) { cudaGraphDestroy(g->graph); g->graph = NULL; }
  // the arena goes when the last handle into it has been released (see nxc_free); at once if none is left
  if (g->live.size() <= g->internal) graph_forget(ctx, g);
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
