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

#include "nxc_common.cuh"

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
  return NXC_OK;
}

extern "C" nxc_status nxc_dist_finalize(nxc_ctx *ctx) {
  if (ctx->nccl_comm) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
    g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = NULL;
  }
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

extern "C" nxc_status nxc_allreduce(nxc_ctx *ctx, void *buf, int64_t count, int dtype, int op) {
  return allreduce_on(ctx, buf, count, dtype, op, ctx->stream);
}
extern "C" nxc_status nxc_allreduce_async(nxc_ctx *ctx, void *buf, int64_t count, int dtype, int op) {
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
  if (!ctx->nccl_comm) { snprintf(ctx->err, sizeof ctx->err, "%s: nxc_dist_init not called", NXC_ERR_NCCL); return NXC_ERR_NCCL; }
  ncclResult_t r = g_nccl.AllGather(send, recv, (size_t)bytes_per_rank, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream);
  if (r) return nccl_fail(ctx, r, "ncclAllGather");
  ctx->launches++;
  return NXC_OK;
}
