// nxc_matmul_tc.cu -- the tensor-core GEMM: tcgen05.mma with the accumulator in
// TMEM, operands staged in shared memory by TMA, one persistent CTA per SM.
//
// Serves bf16 and f16 (f32 accumulate, the reference's "f16/bf16 compute in float"
// rule, nx_c_matmul.c:12-28) and f32 in the opt-in tf32 mode. It replaces the
// reference's pack -> microkernel -> store pipeline (nx_c_matmul.c:363-454,
// 531-587) for those dtypes: "packing through arbitrary strides" becomes a TMA
// tensor map over the view, so a transposed operand (what Rune's backward
// produces, reverse.ml:585-654) is consumed in place:
//
//   operand view            TMA box (inner x outer)      UMMA smem layout
//   A[m,k], a_cs == 1       K(128 B) x 128 rows          K-major,  SWIZZLE_128B
//   A[m,k], a_rs == 1       M(128 B) x BLOCK_K rows, xN   MN-major, SWIZZLE_128B
//   B[k,n], b_rs == 1       K(128 B) x 256 rows          K-major
//   B[k,n], b_cs == 1       N(128 B) x BLOCK_K rows, xN   MN-major
// An MN-major 32-bit operand uses the 32-byte-atom flavour of the 128-byte swizzle on both
// the TMA and the descriptor side (layout type 1, 4-row groups): the only swizzled MN-major
// layout tcgen05 accepts for tf32 -- the plain SWIZZLE_128B one silently yields zeros.
//
// Two instantiations of one kernel:
//   PAIR = false  one CTA per tile, 128x256, tcgen05.mma.cta_group::1 (small M, fallback)
//   PAIR = true   a 2-CTA cluster (the two SMs of a TPC) per 256x256 tile,
//                 tcgen05.mma.cta_group::2 issued by the even CTA: each CTA stages its own
//                 128 rows of A and HALF of B (128 of the 256 columns), so the bytes staged
//                 per flop drop by a third and 7 stages (instead of 4) fit in shared memory;
//                 the accumulator rows 0-127 / 128-255 land in each CTA's own TMEM.
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 MMA issuer (one elected thread) and
// TMEM owner, warps 2-5 epilogue (tcgen05.ld -> convert -> 16-byte global stores).
// Three pipelines: smem full/empty (4 stages x 48 KB), TMEM full/empty (two
// 128x256 f32 accumulators = all 512 columns, so the epilogue of tile i overlaps
// the MMAs of tile i+1), and a static persistent tile schedule that walks N
// within groups of 8 M-blocks so concurrently resident tiles share A and B in L2.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "nxc_matmul.cuh"

namespace {

constexpr int BLOCK_M = 128;    // rows of A (and of the accumulator) one CTA owns
constexpr int BLOCK_N = 256;
constexpr int ROW_BYTES = 128;  // one swizzle row: BLOCK_K * esize
constexpr int A_STAGE_BYTES = BLOCK_M * ROW_BYTES;  // 16 KB
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
template <bool PAIR, bool WIDE = false> struct Cfg {
  static constexpr int TILE_M = PAIR ? 2 * BLOCK_M : BLOCK_M;  // rows of C per tile
  // WIDE (PAIR only): a 256 x 512 tile -- two 256-column halves that share the staged A rows, so a
  // k-block moves 96 KB for twice the flops of the 64 KB a 256 x 256 tile moves (170 instead of 128
  // flop per L2 byte: the pair kernel is bound by L2 -> SM bandwidth). The price:
  // the accumulator takes all 512 TMEM columns, so the epilogue no longer overlaps the next tile.
  static constexpr int TILE_N = WIDE ? 2 * BLOCK_N : BLOCK_N;
  static constexpr int N_HALVES = WIDE ? 2 : 1;
  static constexpr int NUM_ACC = WIDE ? 1 : 2;
  static constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;  // rows of B per CTA and per half
  static constexpr int B_HALF_BYTES = B_ROWS * ROW_BYTES;      // 16 / 32 KB
  static constexpr int B_STAGE_BYTES = N_HALVES * B_HALF_BYTES;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = PAIR ? (WIDE ? 4 : 7) : 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct TcParams {
  int64_t m, n, k;
  int64_t nbatch;
  int64_t c_rs, c_bs;     // element strides of C (row, batch); column stride is 1
  int num_m, num_n, num_kb;
  int esize;              // operand element bytes (2 or 4)
  int block_k;            // elements per k-block = 128 / esize
  int a_mn, b_mn;         // operand is MN-major
  int a_batched, b_batched;
  uint32_t idesc;
  int out_f32;            // C is f32 (tf32 mode) else 16-bit
  int out_bf16;           // 16-bit flavour
  int vec_store;          // rows of C are 16-byte aligned
  int tile_n;             // columns of C per tile: 256, 128 or 64 (the UMMA N; smem / TMEM strides stay those of 256)
  int b_rows;             // rows of B one CTA stages per k-block and per half: tile_n, or tile_n / 2 for a pair
  int group;              // M-blocks per rasterisation group
  int kskew;              // concurrent tiles start their K loop up to kskew-1 blocks apart (1 = in lockstep)
  int prefetch;           // k-blocks ahead of the staged one whose boxes are prefetched into L2 (0 = none)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LAB_DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "LAB_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// the same box, fetched into L2 only (no shared memory, no barrier)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// 64-bit shared-memory matrix descriptor (SWIZZLE_128B, sm_100 version bit)
// layout: 2 = SWIZZLE_128B (16-byte atoms), 1 = SWIZZLE_128B with 32-byte atoms -- the only
// swizzled layout tcgen05 accepts for an MN-major 32-bit (tf32) operand
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint64_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version for Blackwell
  d |= layout << 61;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// ---- 2-CTA (cta_group::2) variants ----------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p`'s counterpart in the even (leader) CTA of the pair
__device__ __forceinline__ uint32_t leader_addr(const void *p, uint32_t leader_rank = 0) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(leader_rank));
  return r;
}
// one TMA box delivered to the same CTA-relative offset of every CTA in `mask`; each destination's
// OWN barrier (same offset) receives the bytes
__device__ __forceinline__ void tma_load_3d_mcast(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, int c2,
                                                  uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "h"(mask)
      : "memory");
}
// both CTAs of the pair load their own slice; the bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap *map, uint32_t leader_bar, void *dst, int c0, int c1,
                                                 int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(leader_bar), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs once the pair's MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar, uint16_t mask = 3) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LAB_WAITC:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LAB_DONEC;\n\t"
      "bra LAB_WAITC;\n\t"
      "LAB_DONEC:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint16_t f32_to_bf16_rn(float f) {
  uint32_t b = __float_as_uint(f);
  if ((b & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((b >> 16) | 0x0040u);
  return (uint16_t)((b + (((b >> 16) & 1u) + 0x7FFFu)) >> 16);
}
__device__ __forceinline__ uint16_t f32_to_f16_rn(float f) {
  uint32_t b = __float_as_uint(f);
  if ((b & 0x7FFFFFFFu) > 0x7F800000u) {
    uint16_t r = (uint16_t)(0x7C00u + ((b & 0x007FFFFFu) >> 13));
    r += (r == 0x7C00u);
    return (uint16_t)(((b & 0x80000000u) >> 16) + r);
  }
  return __half_as_ushort(__float2half_rn(f));
}

__device__ __forceinline__ void tile_coords(const TcParams &p, int64_t t, int &bi, int &mb, int &nb) {
  const int64_t per_batch = (int64_t)p.num_m * p.num_n;
  bi = (int)(t / per_batch);
  int r = (int)(t - (int64_t)bi * per_batch);
  const int GROUP = p.group;
  const int per_group = GROUP * p.num_n;
  const int g = r / per_group;
  const int first_m = g * GROUP;
  const int gsz = (p.num_m - first_m) < GROUP ? (p.num_m - first_m) : GROUP;
  const int rr = r - g * per_group;
  mb = first_m + rr % gsz;
  nb = rr / gsz;
}

// QUAD (PAIR only): a cluster of FOUR CTAs = two pairs working on the tiles (mb, nb) and (mb + 1, nb).
// They need the same B tile, and the pair kernel is bound by operand delivery (halving the tile
// width, +50 % bytes per flop, costs 30 %: profiles/mm_knobs_r02.json), so B is fetched ONCE for
// both: every CTA issues one quarter of the 256-column B slab as a TMA multicast to itself and to
// its sibling in the other pair (ranks r and r ^ 2), plus its own 128 rows of A -- 24 KB of L2
// reads per CTA and k-block instead of 32. Barrier protocol:
//   full[s]   per CTA, one arrival (its own expect_tx) + 32 KB: 16 KB A, 8 KB B issued by itself,
//             8 KB B issued by the sibling. Plain (cta_group::1) TMA semantics: the bytes count on
//             the barrier of the CTA they land in.
//   pfull[s]  on each pair leader: the odd CTA's otherwise idle warp 1 waits for ITS full[s] and
//             forwards one arrival, so the leader issues the pair's MMAs once both halves are in.
//   empty[s]  per CTA, TWO arrivals: the tcgen05.commit of BOTH leaders, multicast to all four CTAs
//             -- a producer's multicast writes into the other pair's shared memory, so a slot is
//             free only when both pairs' MMAs that read it have retired.
// TMEM, the epilogue and the pair-internal tfull / tempty handshake are those of PAIR.
template <bool PAIR, bool WIDE, bool QUAD = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
nxc_mm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 void *__restrict__ Cout, const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  typedef Cfg<PAIR, WIDE> C;
  constexpr int STAGES = C::STAGES, B_STAGE_BYTES = C::B_STAGE_BYTES, STAGE_BYTES = C::STAGE_BYTES;
  constexpr int NUM_ACC = C::NUM_ACC;
  uint8_t *smem_a = smem;
  uint8_t *smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t *bars = (uint64_t *)(smem + STAGES * STAGE_BYTES);
  uint64_t *full = bars, *empty = bars + STAGES, *tfull = bars + 2 * STAGES, *tempty = bars + 2 * STAGES + 2;
  uint64_t *pfull = bars + 2 * STAGES + 4;  // QUAD only
  uint32_t *tmem_slot = (uint32_t *)(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;   // rank in the cluster: 0-1 (PAIR) or 0-3 (QUAD)
  const uint32_t rank = crank & 1u;                        // role in the pair: 0 = leader (issues the MMAs)
  const uint32_t pairid = QUAD ? (crank >> 1) : 0u;        // which of the cluster's two pairs
  const uint32_t lrank = crank & ~1u;                      // cluster rank of this CTA's pair leader
  const int64_t worker = QUAD ? (blockIdx.x >> 2) : PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int64_t nworkers = QUAD ? (gridDim.x >> 2) : PAIR ? (gridDim.x >> 1) : gridDim.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], QUAD ? 2 : 1); mbar_init(&pfull[s], 1); }
    // tempty is only waited on by the leader's MMA warp: 4 epilogue warps of each CTA arrive on it
    for (int s = 0; s < 2; s++) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], PAIR ? 8 : 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // the same warp of both CTAs takes part in a cta_group::2 allocation
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) cluster_sync_all();  // the peer's barriers must be initialised before anything arrives on them
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // QUAD: the work unit is two M-blocks of one N-block; p.num_m then counts UNITS (the host halves it)
  const int64_t total_tiles = (int64_t)p.num_m * p.num_n * p.nbatch;
  // WIDE keeps its compile-time 2 x 256 columns; otherwise the tile width is a launch parameter
  const int tile_n = WIDE ? C::TILE_N : p.tile_n;
  const int b_rows = WIDE ? C::B_ROWS : p.b_rows;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int elems_per_row = ROW_BYTES / p.esize;  // elements in one 128-byte swizzle row
      for (int64_t t = worker; t < total_tiles; t += nworkers) {
        int bi, mb, nb;
        tile_coords(p, t, bi, mb, nb);
        if (QUAD) mb = 2 * mb + (int)pairid;
        const int m0 = mb * C::TILE_M + (int)rank * BLOCK_M, n0 = nb * tile_n + (int)rank * b_rows;
        const int ba = p.a_batched ? bi : 0, bb = p.b_batched ? bi : 0;
        // Tiles that run concurrently walk K in lockstep otherwise: with a power-of-two row pitch
        // (K = 16384 bf16: 32 KB) every CTA then asks for the same 128-byte column of its rows,
        // which lands on a handful of L2 slices / HBM channels. Starting each tile a few k-blocks
        // apart spreads the columns in flight; the window stays small so panels shared through
        // L2 are still hit while hot. The accumulation order is a rotation of 0..num_kb-1.
        const int kb0 = (int)(t % p.kskew);
        for (int kbi = 0; kbi < p.num_kb; kbi++) {
          const int kb = kbi + kb0 < p.num_kb ? kbi + kb0 : kbi + kb0 - p.num_kb;
          mbar_wait(&empty[stage], phase ^ 1);
          // PAIR: one arrival (the leader's) and the bytes of BOTH CTAs complete the leader's barrier.
          // QUAD: every CTA counts the bytes that land in ITS shared memory on its own barrier.
          if (QUAD) mbar_expect_tx(&full[stage], A_STAGE_BYTES + b_rows * ROW_BYTES);
          else if (rank == 0) mbar_expect_tx(&full[stage], (PAIR ? 2 : 1) * (A_STAGE_BYTES + C::N_HALVES * b_rows * ROW_BYTES));
          const uint32_t lbar = (PAIR && !QUAD) ? leader_addr(&full[stage]) : 0u;
          const int k0 = kb * p.block_k;
          uint8_t *sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t *sb = smem_b + stage * B_STAGE_BYTES;
          auto load = [&](const CUtensorMap *map, void *dst, int c0, int c1, int c2) {
            if (PAIR && !QUAD) tma_load_3d_pair(map, lbar, dst, c0, c1, c2);
            else tma_load_3d(map, &full[stage], dst, c0, c1, c2);
          };
          if (!p.a_mn) {
            load(&map_a, sa, k0, m0, ba);
          } else {
            const int box_bytes = p.block_k * ROW_BYTES;
            for (int j = 0; j < BLOCK_M / elems_per_row; j++)
              load(&map_a, sa + j * box_bytes, m0 + j * elems_per_row, k0, ba);
          }
          if (QUAD) {
            // my quarter of the B slab: half of this CTA's b_rows rows, delivered to this CTA and to
            // its sibling in the other pair (cluster ranks rank and rank + 2)
            const uint16_t mask = (uint16_t)((1u << rank) | (1u << (rank + 2)));
            const int half = b_rows / 2;
            if (!p.b_mn) {
              tma_load_3d_mcast(&map_b, &full[stage], sb + pairid * half * ROW_BYTES, k0, n0 + (int)pairid * half, bb, mask);
            } else {
              const int box_bytes = p.block_k * ROW_BYTES, boxes = half / elems_per_row;
              for (int j = 0; j < boxes; j++) {
                const int jj = (int)pairid * boxes + j;
                tma_load_3d_mcast(&map_b, &full[stage], sb + jj * box_bytes, n0 + jj * elems_per_row, k0, bb, mask);
              }
            }
          } else
#pragma unroll
          for (int h = 0; h < C::N_HALVES; h++) {  // WIDE: this CTA's 128 columns of each 256-column half
            uint8_t *sbh = sb + h * C::B_HALF_BYTES;
            const int nh = n0 + h * BLOCK_N;
            if (!p.b_mn) {
              load(&map_b, sbh, k0, nh, bb);
            } else {
              const int box_bytes = p.block_k * ROW_BYTES;
              for (int j = 0; j < b_rows / elems_per_row; j++)
                load(&map_b, sbh + j * box_bytes, nh + j * elems_per_row, k0, bb);
            }
          }
          // A stage is usable only when ALL its boxes have landed, and with a 78 % L2 hit rate most
          // stages contain at least one box that comes from DRAM: ask L2 for the boxes of a later
          // k-block now, so that the staging loads hit.
          if (p.prefetch > 0 && kbi + p.prefetch < p.num_kb) {
            const int kbp = kbi + p.prefetch + kb0 < p.num_kb ? kbi + p.prefetch + kb0 : kbi + p.prefetch + kb0 - p.num_kb;
            const int kp0 = kbp * p.block_k;
            if (!p.a_mn) tma_prefetch_3d(&map_a, kp0, m0, ba);
            else for (int j = 0; j < BLOCK_M / elems_per_row; j++) tma_prefetch_3d(&map_a, m0 + j * elems_per_row, kp0, ba);
            if (!QUAD) {
              for (int h = 0; h < C::N_HALVES; h++) {
                const int nh = n0 + h * BLOCK_N;
                if (!p.b_mn) tma_prefetch_3d(&map_b, kp0, nh, bb);
                else for (int j = 0; j < b_rows / elems_per_row; j++) tma_prefetch_3d(&map_b, nh + j * elems_per_row, kp0, bb);
              }
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (QUAD && warp == 1 && rank == 1) {
    // ===== QUAD, odd CTA: forward "my half of the stage has landed" to the pair leader =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = worker; t < total_tiles; t += nworkers)
        for (int kb = 0; kb < p.num_kb; kb++) {
          mbar_wait(&full[stage], phase);
          mbar_arrive_cluster(leader_addr(&pfull[stage], lrank));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (the leader CTA only when PAIR) =====
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t box_bytes = (uint32_t)p.block_k * ROW_BYTES;
    // K-major: 8-row groups 1024 B apart, a k-step is 32 B inside the swizzle row.
    // MN-major: atoms of 128 B x 8 k-rows, next MN atom one TMA box away, a k-step is
    // (32 / esize) k-rows = that many 128-byte rows.
    // MN-major tf32: the 32-byte-atom swizzle repeats every 4 k-rows (512 B), not 8
    const bool a32 = p.a_mn && p.esize == 4, b32 = p.b_mn && p.esize == 4;
    const uint32_t a_lbo = p.a_mn ? box_bytes : 16, a_sbo = a32 ? 512 : 1024;
    const uint32_t b_lbo = p.b_mn ? box_bytes : 16, b_sbo = b32 ? 512 : 1024;
    const uint64_t a_lay = a32 ? 1 : 2, b_lay = b32 ? 1 : 2;
    const uint32_t kstep_rows = 32 / p.esize;
    const uint32_t a_kstep = p.a_mn ? kstep_rows * ROW_BYTES : 32;
    const uint32_t b_kstep = p.b_mn ? kstep_rows * ROW_BYTES : 32;
    for (int64_t t = worker; t < total_tiles; t += nworkers) {
      if (PAIR) mbar_wait_cluster(&tempty[as], aphase ^ 1);
      else mbar_wait(&tempty[as], aphase ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_base + (uint32_t)as * BLOCK_N;  // WIDE: as == 0, both halves side by side
      for (int kb = 0; kb < p.num_kb; kb++) {
        mbar_wait(&full[stage], phase);
        if (QUAD) mbar_wait_cluster(&pfull[stage], phase);  // the odd CTA's half
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint64_t ad = make_desc(sa + j * a_kstep, a_lbo, a_sbo, a_lay);
            const uint32_t acc = (kb > 0 || j > 0) ? 1u : 0u;
#pragma unroll
            for (int h = 0; h < C::N_HALVES; h++) {
              const uint64_t bd = make_desc(sb + h * C::B_HALF_BYTES + j * b_kstep, b_lbo, b_sbo, b_lay);
              const uint32_t td = tmem_d + (uint32_t)h * BLOCK_N;
              if (PAIR) {
                if (p.esize == 2) umma_f16_pair(td, ad, bd, p.idesc, acc);
                else umma_tf32_pair(td, ad, bd, p.idesc, acc);
              } else {
                if (p.esize == 2) umma_f16(td, ad, bd, p.idesc, acc);
                else umma_tf32(td, ad, bd, p.idesc, acc);
              }
            }
          }
        }
        __syncwarp();
        if (elect_one()) {
          if (PAIR) {  // both CTAs' producers / epilogues are released (QUAD: the producers of all four)
            umma_commit_pair(&empty[stage], QUAD ? (uint16_t)15 : (uint16_t)3);
            if (kb == p.num_kb - 1) umma_commit_pair(&tfull[as], (uint16_t)(3u << lrank));
          } else {
            umma_commit(&empty[stage]);  // frees the smem slot when these MMAs retire
            if (kb == p.num_kb - 1) umma_commit(&tfull[as]);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (p.num_kb == 0) {  // k == 0: nothing accumulates; the epilogue writes zeros
        if (elect_one()) { if (PAIR) umma_commit_pair(&tfull[as], (uint16_t)(3u << lrank)); else umma_commit(&tfull[as]); }
        __syncwarp();
      }
      if (++as == NUM_ACC) { as = 0; aphase ^= 1; }
    }
  } else if (warp >= 2) {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    int as = 0;
    uint32_t aphase = 0;
    for (int64_t t = worker; t < total_tiles; t += nworkers) {
      int bi, mb, nb;
      tile_coords(p, t, bi, mb, nb);
      if (QUAD) mb = 2 * mb + (int)pairid;
      mbar_wait(&tfull[as], aphase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t row = (int64_t)mb * C::TILE_M + (int64_t)rank * BLOCK_M + q * 32 + lane;
      const int64_t col0 = (int64_t)nb * tile_n;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * BLOCK_N;
      const int64_t crow = (int64_t)bi * p.c_bs + row * p.c_rs;
#pragma unroll 1
      for (int c = 0; c < tile_n / 32; c++) {
        uint32_t v[32];
        if (p.num_kb > 0) {
          tmem_ld32(taddr + c * 32, v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] = 0;
        }
        const int64_t col = col0 + c * 32;
        if (row < p.m && col < p.n) {
          if (p.out_f32) {
            float *dst = (float *)Cout + crow + col;
            if (p.vec_store && col + 32 <= p.n) {
#pragma unroll
              for (int i = 0; i < 8; i++)
                *(uint4 *)(dst + 4 * i) = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
              for (int i = 0; i < 32 && col + i < p.n; i++) dst[i] = __uint_as_float(v[i]);
            }
          } else {
            uint16_t h[32];
#pragma unroll
            for (int i = 0; i < 32; i++)
              h[i] = p.out_bf16 ? f32_to_bf16_rn(__uint_as_float(v[i])) : f32_to_f16_rn(__uint_as_float(v[i]));
            uint16_t *dst = (uint16_t *)Cout + crow + col;
            if (p.vec_store && col + 32 <= p.n) {
#pragma unroll
              for (int i = 0; i < 4; i++) {
                uint4 w;
                w.x = h[8 * i] | ((uint32_t)h[8 * i + 1] << 16);
                w.y = h[8 * i + 2] | ((uint32_t)h[8 * i + 3] << 16);
                w.z = h[8 * i + 4] | ((uint32_t)h[8 * i + 5] << 16);
                w.w = h[8 * i + 6] | ((uint32_t)h[8 * i + 7] << 16);
                *(uint4 *)(dst + 8 * i) = w;
              }
            } else {
              for (int i = 0; i < 32 && col + i < p.n; i++) dst[i] = h[i];
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(leader_addr(&tempty[as], lrank));
        else mbar_arrive(&tempty[as]);
      }
      if (++as == NUM_ACC) { as = 0; aphase ^= 1; }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  // PAIR: neither CTA may retire while the other can still read its smem or signal its barriers
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                   : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Build the 3-D tensor map (inner, outer, batch) of one operand.
bool encode_operand(nxc_ctx *ctx, CUtensorMap *map, const void *base, CUtensorMapDataType dt, int esize,
                    int64_t inner_extent, int64_t outer_extent, int64_t outer_stride_elems, int64_t nbatch,
                    int64_t batch_stride_elems, int box_outer, bool atom32 = false) {
  EncodeTiledFn fn = (EncodeTiledFn)ctx->encode_tiled;
  if (!fn) return false;
  if (((uintptr_t)base & 15) != 0) return false;
  const int64_t ob = outer_stride_elems * esize, bb = batch_stride_elems * esize;
  if (ob <= 0 || (ob & 15) != 0) return false;
  if (nbatch > 1 && (bb <= 0 || (bb & 15) != 0)) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)inner_extent, (cuuint64_t)outer_extent, (cuuint64_t)(nbatch > 1 ? nbatch : 1)};
  cuuint64_t gstr[2] = {(cuuint64_t)ob, (cuuint64_t)(nbatch > 1 ? bb : ob * outer_extent)};
  if ((gstr[1] & 15) != 0) gstr[1] = (gstr[1] + 15) & ~(cuuint64_t)15;
  cuuint32_t box[3] = {(cuuint32_t)(ROW_BYTES / esize), (cuuint32_t)box_outer, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, dt, 3, (void *)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace

nxc_status nxc_matmul_tc(nxc_ctx *ctx, const NxcMatmulProblem &q) {
  if (!ctx->encode_tiled) return NXC_MM_TC_DECLINED;
  const int esize = (q.dt == NXC_F32) ? 4 : 2;
  if (q.k == 0 || q.m >= 0x7FFFFFFFLL || q.n >= 0x7FFFFFFFLL || q.k >= 0x7FFFFFFFLL) return NXC_MM_TC_DECLINED;
  // tiny products are launch/latency bound on a 128x256 tile: leave them to the CUDA-core path
  if (q.m * q.n * q.k < (int64_t)64 * 64 * 64) return NXC_MM_TC_DECLINED;
  if (q.c_cs != 1) return NXC_MM_TC_DECLINED;

  // batch dims must collapse to one stride per operand (or a full broadcast)
  int64_t a_bs = 0, b_bs = 0, c_bs = 0;
  bool a_b = false, b_b = false;
  if (q.nbatch > 1) {
    // walk from the innermost batch dim outwards, requiring composition
    int64_t ext = 1;
    bool first = true;
    for (int i = q.batch_nd - 1; i >= 0; i--) {
      if (q.bshape[i] == 1) continue;
      if (first) { a_bs = q.as_[i]; b_bs = q.bs_[i]; c_bs = q.cs_[i]; first = false; }
      else if (q.as_[i] != a_bs * ext || q.bs_[i] != b_bs * ext || q.cs_[i] != c_bs * ext) return NXC_MM_TC_DECLINED;
      ext *= q.bshape[i];
    }
    a_b = a_bs != 0;
    b_b = b_bs != 0;
  }

  // One 256x256 tile per CTA pair when M fills it; the single-CTA 128x256 kernel otherwise
  // (M <= 128, or a 128-row remainder that would leave half of the pair idle too often).
  const int64_t num_m128 = (q.m + BLOCK_M - 1) / BLOCK_M;
  bool pair = num_m128 >= 2 && ((num_m128 & 1) == 0 || num_m128 >= 8);
  if (const char *f = getenv("NX_CUDA_MM_PAIR")) pair = f[0] == '1';  // test / tuning override
  const int tile_m = pair ? 2 * BLOCK_M : BLOCK_M;
  // 256 x 512 tiles (Cfg<true, true>): an experiment that stays opt-in (NX_CUDA_MM_WIDE=1). It moves
  // a quarter fewer L2 bytes per flop and is bit-identical to the 256 x 256 kernel, but measured
  // SLOWER (8192^3: 1284 vs 1395 TFLOP/s; 16384^3 equal): 4 smem stages instead of 7 and an epilogue
  // that no longer overlaps cost more than the saved L2 traffic buys -- so the 256 x 256 kernel's
  // 77 % tensor-pipe activity is not an L2-bandwidth limit (profiles/mm_wide_r01.json).
  bool wide = false;
  if (const char *f = getenv("NX_CUDA_MM_WIDE")) wide = pair && f[0] == '1' && q.n % (2 * BLOCK_N) == 0;
  // Tile width. 256 columns give the most flops per staged byte and are the choice whenever the grid
  // fills the machine; a narrow or small product takes 128 or 64 so that (a) a 64-column C (the
  // head dim of every attention gradient) does not pay for 256, and (b) a product of a few tiles
  // spreads over more SMs (512^3 .. 2048^3 are latency-bound: 4 / 16 / 64 tiles of 256 x 256).
  // An N-major B is staged in boxes of one 128-byte row of N, so a CTA's share must hold whole boxes.
  int tile_n = wide ? 2 * BLOCK_N : BLOCK_N;
  if (!wide) {
    const int64_t workers = pair ? ctx->sm_count / 2 : ctx->sm_count;
    const int64_t num_m_t = (q.m + tile_m - 1) / tile_m;
    const bool b_n_major = !(q.b_rs == 1 || q.k == 1);
    const int epr = ROW_BYTES / esize;
    auto allowed = [&](int tn) { const int br = pair ? tn / 2 : tn; return !b_n_major || br % epr == 0; };
    auto tiles = [&](int tn) { return num_m_t * ((q.n + tn - 1) / tn) * q.nbatch; };
    // waves x (columns per tile + a fixed per-tile cost worth ~32 columns); a narrower tile must win
    // by 15 % to be taken, since it moves more bytes per flop
    auto cost = [&](int tn) { return (double)((tiles(tn) + workers - 1) / workers) * (double)(tn + 32); };
    double best = cost(tile_n);
    for (int tn = 128; tn >= 64; tn >>= 1) {
      if (!allowed(tn)) break;
      const double c = cost(tn);
      if (c < 0.85 * best) { best = c; tile_n = tn; }
    }
    if (const char *f = getenv("NX_CUDA_MM_TILE_N")) { const int v = atoi(f); if ((v == 64 || v == 128 || v == 256) && allowed(v)) tile_n = v; }
  }
  const int b_rows = wide ? (pair ? BLOCK_N / 2 : BLOCK_N) : (pair ? tile_n / 2 : tile_n);
  // Four-CTA clusters sharing B by multicast (see the kernel): products that fill the machine with
  // full-width tiles -- that is where operand delivery binds. Opt-in until measured
  // (NX_CUDA_MM_QUAD=1), off for anything else.
  bool quad = false;
  if (const char *f = getenv("NX_CUDA_MM_QUAD")) {
    const int epr_b = ROW_BYTES / esize;
    const bool b_n_major_q = !(q.b_rs == 1 || q.k == 1);
    quad = f[0] == '1' && pair && !wide && tile_n == BLOCK_N && q.m > 2 * BLOCK_M &&
           (!b_n_major_q || (b_rows / 2) % epr_b == 0);
  }

  TcParams p;
  p.m = q.m; p.n = q.n; p.k = q.k; p.nbatch = q.nbatch;
  p.c_rs = q.c_rs; p.c_bs = c_bs;
  p.esize = esize;
  p.block_k = ROW_BYTES / esize;
  p.num_m = (int)((q.m + tile_m - 1) / tile_m);
  if (quad) p.num_m = (p.num_m + 1) / 2;   // units of two M-blocks (an odd tail computes one block out of range)
  p.num_n = (int)((q.n + tile_n - 1) / tile_n);
  p.num_kb = (int)((q.k + p.block_k - 1) / p.block_k);
  p.a_batched = a_b; p.b_batched = b_b;
  p.tile_n = tile_n; p.b_rows = b_rows;
  p.group = 8;
  if (const char *f = getenv("NX_CUDA_MM_GROUP")) p.group = atoi(f) > 0 ? atoi(f) : 8;
  p.kskew = 1;
  if (const char *f = getenv("NX_CUDA_MM_KSKEW")) p.kskew = atoi(f) > 0 ? atoi(f) : 1;
  if (p.kskew > p.num_kb) p.kskew = p.num_kb > 0 ? p.num_kb : 1;
  p.prefetch = 0;
  if (const char *f = getenv("NX_CUDA_MM_PREFETCH")) p.prefetch = atoi(f) > 0 ? atoi(f) : 0;
  p.out_f32 = (q.dt == NXC_F32);
  p.out_bf16 = (q.dt == NXC_BF16);
  p.vec_store = (((uintptr_t)q.c & 15) == 0) && ((q.c_rs * esize) % 16 == 0) && ((c_bs * esize) % 16 == 0);

  // operand majors
  if (q.a_cs == 1 || q.k == 1) p.a_mn = 0;
  else if (q.a_rs == 1 || q.m == 1) p.a_mn = 1;
  else return NXC_MM_TC_DECLINED;
  if (q.b_rs == 1 || q.k == 1) p.b_mn = 0;
  else if (q.b_cs == 1 || q.n == 1) p.b_mn = 1;
  else return NXC_MM_TC_DECLINED;
  // a K-major operand needs a real row stride; degenerate extents fall back
  if ((q.k == 1 && q.a_cs != 1) || (q.k == 1 && q.b_rs != 1)) return NXC_MM_TC_DECLINED;
  if ((p.a_mn && q.a_rs != 1) || (p.b_mn && q.b_cs != 1)) return NXC_MM_TC_DECLINED;

  const CUtensorMapDataType tdt = q.dt == NXC_BF16  ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                  : q.dt == NXC_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                    : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUtensorMap map_a, map_b;
  bool ok;
  if (!p.a_mn) ok = encode_operand(ctx, &map_a, q.a, tdt, esize, q.k, q.m, q.a_rs, a_b ? q.nbatch : 1, a_bs, BLOCK_M);
  else ok = encode_operand(ctx, &map_a, q.a, tdt, esize, q.m, q.k, q.a_cs, a_b ? q.nbatch : 1, a_bs, p.block_k, esize == 4);
  if (!ok) return NXC_MM_TC_DECLINED;
  if (!p.b_mn) ok = encode_operand(ctx, &map_b, q.b, tdt, esize, q.k, q.n, q.b_cs, b_b ? q.nbatch : 1, b_bs, quad ? b_rows / 2 : b_rows);
  else ok = encode_operand(ctx, &map_b, q.b, tdt, esize, q.n, q.k, q.b_rs, b_b ? q.nbatch : 1, b_bs, p.block_k, esize == 4);
  if (!ok) return NXC_MM_TC_DECLINED;

  // instruction descriptor: D = f32, A/B format, majors, N >> 3, M >> 4 (M = 256 across the pair)
  const uint32_t fmt = q.dt == NXC_BF16 ? 1u : q.dt == NXC_F16 ? 0u : 2u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
            ((uint32_t)((wide ? BLOCK_N : tile_n) >> 3) << 17) | ((uint32_t)(tile_m >> 4) << 24);

  if (!ctx->mm_attr_set) {  // per context = per device; a function attribute is device state
    NXC_CUDA_TRY(ctx, cudaFuncSetAttribute(nxc_mm_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg<false>::SMEM_BYTES));
    NXC_CUDA_TRY(ctx, cudaFuncSetAttribute(nxc_mm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg<true>::SMEM_BYTES));
    NXC_CUDA_TRY(ctx, cudaFuncSetAttribute(nxc_mm_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg<true, true>::SMEM_BYTES));
    NXC_CUDA_TRY(ctx, cudaFuncSetAttribute(nxc_mm_tc_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg<true>::SMEM_BYTES));
    ctx->mm_attr_set = 1;
  }
  const int64_t tiles = (int64_t)p.num_m * p.num_n * p.nbatch;
  if (!pair) {
    const int grid = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
    nxc_mm_tc_kernel<false, false><<<grid, NUM_THREADS, Cfg<false>::SMEM_BYTES, ctx->stream>>>(map_a, map_b, (void *)q.c, p);
  } else if (quad) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = Cfg<true>::SMEM_BYTES;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    static int quad_clusters = 0;       // one device per process (one process per GPU)
    if (quad_clusters == 0) {   // how many 4-CTA clusters of this kernel the device runs at once
      int nc = 0;
      cfg.gridDim = dim3((unsigned)(4 * (ctx->sm_count / 4)));
      if (cudaOccupancyMaxActiveClusters(&nc, nxc_mm_tc_kernel<true, false, true>, &cfg) != cudaSuccess || nc < 1) {
        cudaGetLastError();
        nc = ctx->sm_count / 4 - 2;
      }
      quad_clusters = nc;
    }
    const int64_t clusters = quad_clusters;
    cfg.gridDim = dim3((unsigned)(4 * (tiles < clusters ? tiles : clusters)));
    void *out = (void *)q.c;
    cudaError_t e = cudaLaunchKernelEx(&cfg, nxc_mm_tc_kernel<true, false, true>, map_a, map_b, out, p);
    if (e != cudaSuccess) return nxc_cuda_fail(ctx, e, "cluster launch (4 CTAs)");
  } else {
    const int64_t pairs = ctx->sm_count / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)(2 * (tiles < pairs ? tiles : pairs)));
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = wide ? Cfg<true, true>::SMEM_BYTES : Cfg<true>::SMEM_BYTES;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    void *out = (void *)q.c;
    cudaError_t e = wide ? cudaLaunchKernelEx(&cfg, nxc_mm_tc_kernel<true, true>, map_a, map_b, out, p)
                         : cudaLaunchKernelEx(&cfg, nxc_mm_tc_kernel<true, false>, map_a, map_b, out, p);
    if (e != cudaSuccess) return nxc_cuda_fail(ctx, e, "cluster launch");
  }
  NXC_LAUNCH_CHECK(ctx);
  return NXC_OK;
}
