// nxc_dist.cuh -- the peer-memory mailbox protocol shared by the exchange kernels
// (nxc_dist.cu: all-gather; nxc_dist_fold.cu: all-reduce and the arg-reduce finish).
//
// Every rank owns a MAILBOX (plain cudaMalloc, exported with CUDA IPC and mapped by every peer at
// nxc_dist_init). It holds, per parity of the exchange EPOCH and per source rank, one 256 KiB
// slot cut into 16 KiB chunks, and one flag word per (parity, source, chunk). A sender stores
// its chunk straight into the receiver's slot over NVLink and then releases the flag there with
// the epoch number; a receiver spins on the flag in its OWN memory. A slot is rewritten at epoch
// e + 2, which a sender reaches only after it saw this rank's epoch e + 1 flags, i.e. after this
// rank finished reading epoch e.
//
// The epoch lives in DEVICE memory (state[0], bumped by the last CTA of each exchange kernel), not
// in a host counter passed as a kernel argument: an exchange is then a pure function of device
// state, so a captured step (nxc_capture_begin) replays it correctly any number of times, and
// eager and replayed exchanges may be mixed as long as every rank issues the same sequence.
#pragma once

#include "nxc_map.cuh"

#define NXC_P2P_MAX_WORLD 16
#define NXC_P2P_SLOT_BYTES ((size_t)256 << 10)
#define NXC_P2P_CHUNK_BYTES ((size_t)16 << 10)
#define NXC_P2P_MAX_CHUNKS ((int)(NXC_P2P_SLOT_BYTES / NXC_P2P_CHUNK_BYTES))
// outputs per chunk of the arg-reduce finish: values (up to 8 bytes) then int32 indices
#define NXC_P2P_ARG_PER_CHUNK 1024

struct nxc_p2p {
  int world, rank;
  char *local;                      // this rank's mailbox
  char *peer[NXC_P2P_MAX_WORLD];    // every rank's mailbox as mapped here (peer[rank] == local)
  uint32_t *state;                  // device: [0] epoch of the last finished exchange, [1] CTAs done
};

// what every exchange kernel gets
struct NxcP2P {
  char *peer[NXC_P2P_MAX_WORLD];
  int world, rank;
  uint32_t *state;
  int *status;                      // the context's status page (mapped host memory)
};

static __host__ __device__ inline size_t nxc_p2p_data_bytes(int world) { return 2 * (size_t)world * NXC_P2P_SLOT_BYTES; }
static inline size_t nxc_p2p_total_bytes(int world) {
  return nxc_p2p_data_bytes(world) + 2 * (size_t)world * NXC_P2P_MAX_CHUNKS * sizeof(uint32_t);
}
static inline NxcP2P nxc_p2p_args(const nxc_ctx *ctx) {
  const nxc_p2p *q = ctx->p2p;
  NxcP2P a;
  for (int r = 0; r < NXC_P2P_MAX_WORLD; r++) a.peer[r] = r < q->world ? q->peer[r] : NULL;
  a.world = q->world; a.rank = q->rank;
  a.state = q->state;
  a.status = ctx->dstatus + NXC_ST_EXCHANGE;
  return a;
}

#ifdef __CUDACC__
__device__ __forceinline__ void nxc_st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t nxc_ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t nxc_globaltimer() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// the epoch this kernel runs as (every CTA reads it before any CTA of the kernel can publish)
__device__ __forceinline__ uint32_t nxc_p2p_epoch(const NxcP2P &a) { return *(volatile uint32_t *)a.state + 1u; }
// slot of `src`'s payload in rank `at`'s mailbox
__device__ __forceinline__ char *nxc_p2p_slot(const NxcP2P &a, int at, uint32_t e, int src) {
  return a.peer[at] + ((size_t)((e & 1u) * a.world + src)) * NXC_P2P_SLOT_BYTES;
}
__device__ __forceinline__ uint32_t *nxc_p2p_flag(const NxcP2P &a, int at, uint32_t e, int src, int chunk) {
  return (uint32_t *)(a.peer[at] + nxc_p2p_data_bytes(a.world)) + ((size_t)((e & 1u) * a.world + src)) * NXC_P2P_MAX_CHUNKS + chunk;
}
// One thread waits for `flag` (in this rank's own mailbox) to show epoch e. 20 s without it means
// a peer died or never issued the matching call: the status page records it (nxc_sync / nxc_d2h
// raise), the caller skips its reads, and the kernel still retires -- never a hang.
__device__ __forceinline__ bool nxc_p2p_wait(const uint32_t *flag, uint32_t e, int *status) {
  const uint64_t t0 = nxc_globaltimer();
  while (nxc_ld_acquire_sys(flag) != e) {
    __nanosleep(64);
    if (nxc_globaltimer() - t0 > 20000000000ull) {
      *(volatile int *)status = 1;
      __threadfence_system();
      return false;
    }
  }
  return true;
}
// Last statement of every exchange kernel, all threads: the CTA that retires last publishes the
// epoch for the next kernel on the stream.
__device__ __forceinline__ void nxc_p2p_finish(const NxcP2P &a, uint32_t e) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(a.state + 1, 1u);
    if (done == gridDim.x * gridDim.y - 1) {
      a.state[1] = 0;
      *(volatile uint32_t *)a.state = e;
      __threadfence();
    }
  }
}
// bytes -> a peer's (or this rank's) slot; 16-byte vectors when both sides allow
__device__ __forceinline__ void nxc_p2p_copy_out(char *dst, const char *src, int64_t len) {
  if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
    const int64_t nv = len >> 4;
    for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) ((uint4 *)dst)[i] = ((const uint4 *)src)[i];
    for (int64_t i = (nv << 4) + threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
  } else {
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
  }
}
// one element out of this rank's mailbox: the bytes were written by a peer over NVLink, so the
// load must not be served from a stale L1 line of an earlier epoch
template <typename S> __device__ __forceinline__ S nxc_p2p_load(const S *p) {
  typedef typename NxcVecT<sizeof(S)>::T V;
  union { V v; S s; } u;
  u.v = __ldcv(reinterpret_cast<const V *>(p));
  return u.s;
}
#endif

// nxc_dist_fold.cu
nxc_status nxc_p2p_allreduce_fused(nxc_ctx *ctx, void *buf, int64_t count, int dtype, int op);
