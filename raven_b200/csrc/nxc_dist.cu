// nxc_dist.cu -- the multi-GPU exchange step: NCCL allreduce / allgather over
// NVLink 5 / NVSwitch, one process per GPU. The
// reference's backend contract has no collective; these entry points are what
// the sharded reduce / argreduce / batch-matmul host paths and a Kaun
// data-parallel gradient hook call after the local kernels.
//
// NCCL is dlopen'ed (the copy already loaded into the process by torch, else
// libnccl.so.2) so libnxcuda.so loads on a box without it and has no link-time
// dependency. Only the handful of symbols used are declared here.
#include <dlfcn.h>

#include "nxc_dist.cuh"

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
// enum values from the public nccl.h (stable across NCCL 2.x)
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5,
       ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8, ncclBfloat16 = 9 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

static struct {
  void *h;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
  const char *(*GetErrorString)(ncclResult_t);
} g_nccl;

static nxc_status nccl_load(nxc_ctx *ctx) {
  if (g_nccl.h) return NXC_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = NULL;
  for (int i = 0; i < 2 && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s: cannot dlopen libnccl.so.2 (%s)", NXC_ERR_NCCL, dlerror());
    return NXC_ERR_NCCL;
  }
#define NXC_SYM(field, name)                                                        \
  *(void **)(&g_nccl.field) = dlsym(h, name);                                       \
  if (!g_nccl.field) {                                                              \
    if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s: missing symbol %s", NXC_ERR_NCCL, name); \
    return NXC_ERR_NCCL;                                                            \
  }
  NXC_SYM(GetUniqueId, "ncclGetUniqueId")
  NXC_SYM(CommInitRank, "ncclCommInitRank")
  NXC_SYM(CommDestroy, "ncclCommDestroy")
  NXC_SYM(AllReduce, "ncclAllReduce")
  NXC_SYM(AllGather, "ncclAllGather")
  NXC_SYM(GetErrorString, "ncclGetErrorString")
#undef NXC_SYM
  g_nccl.h = h;
  return NXC_OK;
}

static void nxc_p2p_setup(nxc_ctx *ctx);
static void nxc_p2p_teardown(nxc_ctx *ctx);

static nxc_status nccl_fail(nxc_ctx *ctx, ncclResult_t r, const char *what) {
  if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s: %s (%s)", NXC_ERR_NCCL, g_nccl.GetErrorString(r), what);
  return NXC_ERR_NCCL;
}

extern "C" nxc_status nxc_dist_unique_id(void *id_out) {
  nxc_status s = nccl_load(NULL);
  if (s) return s;
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r) return NXC_ERR_NCCL;
  memcpy(id_out, &id, NXC_UNIQUE_ID_BYTES);
  return NXC_OK;
}

extern "C" nxc_status nxc_dist_init(nxc_ctx *ctx, int rank, int world, const void *id_128) {
  NXC_TRACE(ctx, "nxc_dist_init");
  nxc_status s = nccl_load(ctx);
  if (s) return s;
  NXC_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id_128, NXC_UNIQUE_ID_BYTES);
  ncclComm_t comm;
  ncclResult_t r = g_nccl.CommInitRank(&comm, world, id, rank);
  if (r) return nccl_fail(ctx, r, "ncclCommInitRank");
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->world = world;
  nxc_p2p_setup(ctx);
  return NXC_OK;
}

extern "C" nxc_status nxc_dist_finalize(nxc_ctx *ctx) {
  if (ctx->nccl_comm) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
    nxc_p2p_teardown(ctx);
    g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = NULL;
  }
  return NXC_OK;
}

// ---- small exchanges over peer memory ----------------------------------------------------
// The exchange step of a sharded reduction moves a few bytes to a few KB per rank; through
// NCCL that is one ~10-20 us kernel per call. Here ONE kernel does the whole all-gather over
// the mailboxes of nxc_dist.cuh: CTA (s, c) stores chunk c of this rank's payload straight into
// peer s's mailbox over NVLink, releases a flag there, then waits for chunk c from rank s in its
// OWN mailbox and copies it to the output. nxc_allreduce and the arg-reduce finish go one step
// further (nxc_dist_fold.cu): the same kernel that receives the partials folds them, in rank
// order on every rank, so all ranks hold bit-identical results (NCCL promises that only per
// algorithm choice). Payloads above the slot size, async (comm-stream) collectives and
// NX_CUDA_P2P=0 use NCCL.
__global__ void __launch_bounds__(256) nxc_p2p_allgather_kernel(const NxcP2P a, int chunks, int64_t bytes,
                                                                const char *__restrict__ send, char *__restrict__ recv) {
  const int s = blockIdx.x / chunks, c = blockIdx.x - s * chunks;
  const uint32_t e = nxc_p2p_epoch(a);
  const int64_t lo = (int64_t)c * (int64_t)NXC_P2P_CHUNK_BYTES;
  int64_t len = bytes - lo;
  if (len > (int64_t)NXC_P2P_CHUNK_BYTES) len = NXC_P2P_CHUNK_BYTES;
  // push: my chunk c -> peer s
  nxc_p2p_copy_out(nxc_p2p_slot(a, s, e, a.rank) + lo, send + lo, len);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) nxc_st_release_sys(nxc_p2p_flag(a, s, e, a.rank, c), e);
  // pull: chunk c of rank s has to land in my mailbox
  __shared__ int ok;
  if (threadIdx.x == 0) ok = nxc_p2p_wait(nxc_p2p_flag(a, a.rank, e, s, c), e, a.status) ? 1 : 0;
  __syncthreads();
  if (ok) {
    const char *src = nxc_p2p_slot(a, a.rank, e, s) + lo;
    char *dst = recv + (int64_t)s * bytes + lo;
    if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
      const int64_t nv = len >> 4;
      for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) ((uint4 *)dst)[i] = __ldcv((const uint4 *)src + i);
      for (int64_t i = (nv << 4) + threadIdx.x; i < len; i += blockDim.x) dst[i] = __ldcv(src + i);
    } else {
      for (int64_t i = threadIdx.x; i < len; i += blockDim.x) dst[i] = __ldcv(src + i);
    }
  }
  nxc_p2p_finish(a, e);
}

static void nxc_p2p_teardown(nxc_ctx *ctx) {
  nxc_p2p *q = ctx->p2p;
  if (!q) return;
  for (int r = 0; r < q->world; r++)
    if (r != q->rank && q->peer[r]) cudaIpcCloseMemHandle(q->peer[r]);
  if (q->local) cudaFree(q->local);
  if (q->state) cudaFree(q->state);
  free(q);
  ctx->p2p = NULL;
  cudaGetLastError();
}

// Collective: every rank calls it right after ncclCommInitRank. Failure at any step (no peer
// access, IPC refused by the container) is not an error: ALL ranks then stay on NCCL -- the
// decision is taken jointly through an allreduce(min) of the local outcome.
static void nxc_p2p_setup(nxc_ctx *ctx) {
  const char *env = getenv("NX_CUDA_P2P");
  int want = !(env && env[0] == '0') && ctx->world > 1 && ctx->world <= NXC_P2P_MAX_WORLD;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  nxc_p2p *q = (nxc_p2p *)calloc(1, sizeof *q);
  int good = want && q != NULL;
  struct Xch { cudaIpcMemHandle_t h; int ok; int pad[3]; };
  Xch *dev = NULL, *host = (Xch *)calloc(ctx->world, sizeof(Xch));
  Xch mine;
  memset(&mine, 0, sizeof mine);
  if (good) {
    q->world = ctx->world; q->rank = ctx->rank;
    good = cudaMalloc(&q->local, nxc_p2p_total_bytes(ctx->world)) == cudaSuccess &&
           cudaMemset(q->local, 0, nxc_p2p_total_bytes(ctx->world)) == cudaSuccess &&
           cudaMalloc(&q->state, 4 * sizeof(uint32_t)) == cudaSuccess && cudaMemset(q->state, 0, 4 * sizeof(uint32_t)) == cudaSuccess &&
           cudaIpcGetMemHandle(&mine.h, q->local) == cudaSuccess;
  }
  mine.ok = good;
  // the handles travel through NCCL (every rank takes part, whatever its local outcome)
  bool xch = host && cudaMalloc(&dev, sizeof(Xch) * ctx->world) == cudaSuccess &&
             cudaMemcpy(dev + ctx->rank, &mine, sizeof mine, cudaMemcpyHostToDevice) == cudaSuccess &&
             g_nccl.AllGather(dev + ctx->rank, dev, sizeof(Xch), ncclUint8, comm, ctx->stream) == 0 &&
             cudaStreamSynchronize(ctx->stream) == cudaSuccess &&
             cudaMemcpy(host, dev, sizeof(Xch) * ctx->world, cudaMemcpyDeviceToHost) == cudaSuccess;
  if (!xch) good = 0;
  for (int r = 0; good && r < ctx->world; r++)
    if (!host[r].ok) good = 0;
  for (int r = 0; good && r < ctx->world; r++) {
    if (r == ctx->rank) { q->peer[r] = q->local; continue; }
    void *m = NULL;
    if (cudaIpcOpenMemHandle(&m, host[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) good = 0;
    q->peer[r] = (char *)m;
  }
  // joint verdict: min over ranks of `good` (mapping can fail on one rank only)
  int *vd = (int *)dev;
  if (xch && cudaMemcpy(vd, &good, sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess &&
      g_nccl.AllReduce(vd, vd, 1, ncclInt32, ncclMin, comm, ctx->stream) == 0 &&
      cudaStreamSynchronize(ctx->stream) == cudaSuccess) {
    int all = 0;
    if (cudaMemcpy(&all, vd, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) all = 0;
    good = good && all;
  } else {
    good = 0;
  }
  cudaGetLastError();
  if (dev) cudaFree(dev);
  free(host);
  ctx->p2p = q;
  if (!good) nxc_p2p_teardown(ctx);
}

extern "C" int nxc_dist_p2p_enabled(nxc_ctx *ctx) { return ctx->p2p != NULL; }

static nxc_status dist_usable(nxc_ctx *ctx) {
  if (!ctx->nccl_comm) { snprintf(ctx->err, sizeof ctx->err, "%s: nxc_dist_init not called", NXC_ERR_NCCL); return NXC_ERR_NCCL; }
  if (ctx->dist_poisoned || (ctx->hstatus && ctx->hstatus[NXC_ST_EXCHANGE])) {
    ctx->dist_poisoned = 1;
    snprintf(ctx->err, sizeof ctx->err, "%s: %s earlier on this communicator", NXC_ERR_NCCL, NXC_ERR_EXCHANGE_TIMEOUT);
    return NXC_ERR_NCCL;
  }
  return NXC_OK;
}

static nxc_status nxc_p2p_allgather(nxc_ctx *ctx, const void *send, void *recv, int64_t bytes) {
  int chunks = (int)((bytes + (int64_t)NXC_P2P_CHUNK_BYTES - 1) / (int64_t)NXC_P2P_CHUNK_BYTES);
  if (chunks < 1) chunks = 1;
  nxc_p2p_allgather_kernel<<<ctx->p2p->world * chunks, 256, 0, ctx->stream>>>(nxc_p2p_args(ctx), chunks, bytes,
                                                                              (const char *)send, (char *)recv);
  NXC_LAUNCH_CHECK(ctx);
  return NXC_OK;
}

static int nccl_dtype(int dt) {
  switch (dt) {
    case NXC_I8: return ncclInt8;
    case NXC_U8: case NXC_BOOL: return ncclUint8;
    case NXC_I32: return ncclInt32;
    case NXC_U32: return ncclUint32;
    case NXC_I64: return ncclInt64;
    case NXC_U64: return ncclUint64;
    case NXC_F16: return ncclFloat16;
    case NXC_F32: return ncclFloat32;
    case NXC_F64: return ncclFloat64;
    case NXC_BF16: return ncclBfloat16;
    default: return -1;
  }
}

static nxc_status allreduce_on(nxc_ctx *ctx, void *buf, int64_t count, int dtype, int op, cudaStream_t stream);

// allgather of the partials (peer memory or NCCL, by size) + the backend's own reduce over the rank
// axis: what serves payloads above the mailbox slot for dtypes / ops NCCL has no reduction for
// (int16, uint16, fp8, complex prod) and keeps float max / min NaN-sticky
static nxc_status allreduce_by_gather(nxc_ctx *ctx, void *buf, int64_t count, int dtype, int op) {
  const int64_t bytes = count * nxc_elem_size(dtype);
  void *g = NULL;
  nxc_status s = nxc_alloc(ctx, (size_t)bytes * ctx->world, &g);
  if (s) return s;
  s = nxc_allgather(ctx, buf, g, bytes);
  if (!s) {
    nxc_tensor in, out;
    memset(&in, 0, sizeof in);
    memset(&out, 0, sizeof out);
    in.data = g; in.dtype = dtype; in.ndim = 2;
    in.shape[0] = ctx->world; in.shape[1] = count; in.strides[0] = count; in.strides[1] = 1;
    out.data = buf; out.dtype = dtype; out.ndim = 1; out.shape[0] = count; out.strides[0] = 1;
    const int axis = 0;
    s = nxc_reduce(ctx, op, &out, &in, &axis, 1);
  }
  nxc_status f = nxc_free(ctx, g);
  return s ? s : f;
}

extern "C" nxc_status nxc_allreduce(nxc_ctx *ctx, void *buf, int64_t count, int dtype, int op) {
  NXC_TRACE(ctx, "nxc_allreduce");
  nxc_status s = dist_usable(ctx);
  if (s) return s;
  if (op < 0 || op > NXC_RMIN || !nxc_valid_dtype(dtype) || nxc_is_packed(dtype)) return NXC_ERR_UNSUPPORTED_DTYPE;
  if (count <= 0) return NXC_OK;
  const int64_t bytes = count * nxc_elem_size(dtype);
  // one kernel: exchange over peer memory + fold in rank order (nxc_dist_fold.cu)
  if (ctx->p2p && (size_t)bytes <= NXC_P2P_SLOT_BYTES) return nxc_p2p_allreduce_fused(ctx, buf, count, dtype, op);
  const bool is_float = (nxc_dtype_class(dtype) & (NXC_CLS_FLOAT | NXC_CLS_COMPLEX)) != 0;
  const bool nccl_can = nccl_dtype(dtype) >= 0 || ((dtype == NXC_C32 || dtype == NXC_C64) && op == NXC_SUM);
  if (!nccl_can || dtype == NXC_BOOL || (is_float && (op == NXC_RMAX || op == NXC_RMIN)))
    return allreduce_by_gather(ctx, buf, count, dtype, op);
  return allreduce_on(ctx, buf, count, dtype, op, ctx->stream);
}
extern "C" nxc_status nxc_allreduce_async(nxc_ctx *ctx, void *buf, int64_t count, int dtype, int op) {
  NXC_TRACE(ctx, "nxc_allreduce_async");
  nxc_status s = nxc_side_streams(ctx);
  if (s) return s;
  NXC_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
  NXC_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_fork, 0));
  return allreduce_on(ctx, buf, count, dtype, op, ctx->comm_stream);
}
extern "C" nxc_status nxc_comm_wait(nxc_ctx *ctx) {
  if (!ctx->comm_stream) return NXC_OK;
  NXC_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
  NXC_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
  return NXC_OK;
}
static nxc_status allreduce_on(nxc_ctx *ctx, void *buf, int64_t count, int dtype, int op, cudaStream_t stream) {
  if (!ctx->nccl_comm) { snprintf(ctx->err, sizeof ctx->err, "%s: nxc_dist_init not called", NXC_ERR_NCCL); return NXC_ERR_NCCL; }
  int nd = nccl_dtype(dtype);
  int64_t n = count;
  // complex sums reduce as 2x the real component count
  if (dtype == NXC_C32 && op == NXC_SUM) { nd = ncclFloat32; n = 2 * count; }
  if (dtype == NXC_C64 && op == NXC_SUM) { nd = ncclFloat64; n = 2 * count; }
  if (nd < 0 || op < 0 || op > NXC_RMIN) return NXC_ERR_UNSUPPORTED_DTYPE;
  ncclResult_t r = g_nccl.AllReduce(buf, buf, (size_t)n, nd, op, (ncclComm_t)ctx->nccl_comm, stream);
  if (r) return nccl_fail(ctx, r, "ncclAllReduce");
  ctx->launches++;
  return NXC_OK;
}

extern "C" nxc_status nxc_allgather(nxc_ctx *ctx, const void *send, void *recv, int64_t bytes_per_rank) {
  NXC_TRACE(ctx, "nxc_allgather");
  nxc_status us = dist_usable(ctx);
  if (us) return us;
  if (ctx->p2p && bytes_per_rank > 0 && (size_t)bytes_per_rank <= NXC_P2P_SLOT_BYTES)
    return nxc_p2p_allgather(ctx, send, recv, bytes_per_rank);
  ncclResult_t r = g_nccl.AllGather(send, recv, (size_t)bytes_per_rank, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream);
  if (r) return nccl_fail(ctx, r, "ncclAllGather");
  ctx->launches++;
  return NXC_OK;
}
