// nxc_map.cuh -- the map (elementwise) launch machinery.
//
// Host side: the iteration plan. K operands sharing one shape are stripped of
// size-1 dims and merged wherever adjacent dims compose on EVERY operand
// (stride[outer] == stride[inner] * shape[inner]) -- the same effect as the
// reference's nx_c_coalesce_map (nx_c_engine.c:542-590), restated for a grid
// instead of a thread pool: there is no run/odometer split, the coalesced plan
// picks one of two kernels.
//
//   flat    every operand is either dense (unit stride over the one coalesced
//           dim) or a full broadcast (all strides 0): 128-bit vector loads and
//           stores, UNROLL independent 16-byte requests in flight per operand
//           per thread. This is the HBM-roofline path (2^28-element arrays).
//   strided anything else: one work item = VW consecutive elements of the
//           innermost coalesced dim (VW > 1 when every operand's inner stride
//           is 0 or 1 and alignment allows -- row/column broadcasts, slices of
//           rows), outer coordinates by fast 32-bit division.
//
// Device side: a kernel op is a struct K with
//   static constexpr int NIN;                 number of inputs (0..3)
//   typedef ... S0;  S1; S2; S3;              storage types (S0 = output)
//   static __device__ S0 run(S1, S2, S3, const P&)   one element
// so kernels see storage types only; converters live in the op.
#pragma once

#include <type_traits>

#include "nxc_common.cuh"

#define NXC_MAX_OPERANDS 4

struct NxcMapPlan {
  int nop;
  int ndim;  // coalesced rank >= 1
  int64_t shape[NXC_MAX_NDIM];
  int64_t stride[NXC_MAX_OPERANDS][NXC_MAX_NDIM];  // ELEMENT strides
  char *base[NXC_MAX_OPERANDS];                    // data + offset*esize
  int64_t total;
};

// Builds the plan; returns NXC_ERR_OUT_ALIASED if the output (operand 0) has a
// zero stride over an extent > 1 on a non-empty tensor (reference:
// nx_c_engine.c:845-850).
nxc_status nxc_map_plan(const nxc_tensor *const *ops, int nop, const int64_t *esize,
                        NxcMapPlan *p);

// Fast division of a 31-bit dividend by a runtime-constant divisor.
struct NxcFastDiv {
  uint32_t d, mul, sh;
};
static inline NxcFastDiv nxc_fastdiv_make(uint32_t d) {
  NxcFastDiv f;
  f.d = d;
  if (d <= 1) { f.mul = 0; f.sh = 0; return f; }
  uint32_t lg = 0;
  while ((1ull << lg) < d) lg++;
  uint32_t p = 31 + lg;
  f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
  f.sh = p - 32;
  return f;
}
__device__ __forceinline__ uint32_t nxc_fastdiv(uint32_t n, const NxcFastDiv &f) {
  return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.sh);
}

template <int NOP>
struct NxcStridedArgs {
  int ndim;                 // outer dims (the inner dim is handled separately)
  uint32_t inner;           // inner extent in work items
  uint32_t bcast_mask;      // bit k: operand k has inner stride 0
  NxcFastDiv inner_div;
  NxcFastDiv div[NXC_MAX_NDIM];
  int64_t stride[NOP][NXC_MAX_NDIM];  // outer strides, elements
  int64_t inner_stride[NOP];          // elements per work item
  int64_t nitems;                     // total work items
  // 64-bit fallback (nitems >= 2^31)
  int64_t shape64[NXC_MAX_NDIM];
  int64_t inner64;
};

// ---- vector load/store helpers ---------------------------------------------------
template <int BYTES> struct NxcVecT;
template <> struct NxcVecT<1> { typedef uint8_t T; };
template <> struct NxcVecT<2> { typedef uint16_t T; };
template <> struct NxcVecT<4> { typedef uint32_t T; };
template <> struct NxcVecT<8> { typedef uint2 T; };
template <> struct NxcVecT<16> { typedef uint4 T; };

// Load N consecutive elements of S from an N*sizeof(S)-aligned address, in
// chunks of at most 16 bytes.
template <typename S, int N>
__device__ __forceinline__ void nxc_load_vec(const S *p, S (&r)[N]) {
  constexpr int BYTES = N * (int)sizeof(S);
  constexpr int CH = BYTES >= 16 ? 16 : BYTES;
  typedef typename NxcVecT<CH>::T V;
  constexpr int NCH = BYTES / CH;
  const V *pv = reinterpret_cast<const V *>(p);
  V *rv = reinterpret_cast<V *>(&r[0]);
#pragma unroll
  for (int i = 0; i < NCH; i++) rv[i] = __ldg(pv + i);
}
template <typename S, int N>
__device__ __forceinline__ void nxc_store_vec(S *p, const S (&r)[N]) {
  constexpr int BYTES = N * (int)sizeof(S);
  constexpr int CH = BYTES >= 16 ? 16 : BYTES;
  typedef typename NxcVecT<CH>::T V;
  constexpr int NCH = BYTES / CH;
  V *pv = reinterpret_cast<V *>(p);
  const V *rv = reinterpret_cast<const V *>(&r[0]);
#pragma unroll
  for (int i = 0; i < NCH; i++) pv[i] = rv[i];
}

template <typename A, typename B> struct NxcMinSize { static constexpr int v = sizeof(A) < sizeof(B) ? sizeof(A) : sizeof(B); };
template <typename A, typename B> struct NxcMaxSize { static constexpr int v = sizeof(A) > sizeof(B) ? sizeof(A) : sizeof(B); };

template <class K> struct NxcKInfo {
  static constexpr int s0 = sizeof(typename K::S0);
  static constexpr int s1 = K::NIN >= 1 ? sizeof(typename K::S1) : s0;
  static constexpr int s2 = K::NIN >= 2 ? sizeof(typename K::S2) : s0;
  static constexpr int s3 = K::NIN >= 3 ? sizeof(typename K::S3) : s0;
  static constexpr int mn01 = s0 < s1 ? s0 : s1, mn23 = s2 < s3 ? s2 : s3;
  static constexpr int mx01 = s0 > s1 ? s0 : s1, mx23 = s2 > s3 ? s2 : s3;
  static constexpr int min_size = mn01 < mn23 ? mn01 : mn23;
  static constexpr int max_size = mx01 > mx23 ? mx01 : mx23;
  // items per thread per step: the smallest type moves 16 bytes
  static constexpr int IPT = 16 / min_size;
  // independent steps in flight: aim at 64 bytes of the largest type per thread
  static constexpr int UNROLL_RAW = 64 / (IPT * max_size);
  static constexpr int UNROLL = UNROLL_RAW < 1 ? 1 : (UNROLL_RAW > 4 ? 4 : UNROLL_RAW);
};

#define NXC_MAP_THREADS 256

// ---- flat kernel -------------------------------------------------------------------
template <class K>
__global__ void __launch_bounds__(NXC_MAP_THREADS)
nxc_map_flat_kernel(typename K::S0 *__restrict__ out, const typename K::S1 *__restrict__ a,
                    const typename K::S2 *__restrict__ b, const typename K::S3 *__restrict__ c,
                    int64_t n, uint32_t bcast, typename K::P prm) {
  typedef typename K::S0 S0; typedef typename K::S1 S1; typedef typename K::S2 S2; typedef typename K::S3 S3;
  constexpr int IPT = NxcKInfo<K>::IPT, UNROLL = NxcKInfo<K>::UNROLL;
  constexpr int64_t TILE = (int64_t)NXC_MAP_THREADS * IPT * UNROLL;
  const int64_t tile0 = (int64_t)blockIdx.x * TILE;
  S1 sa = S1(); S2 sb = S2(); S3 sc = S3();
  if (K::NIN >= 1 && (bcast & 2u)) sa = a[0];
  if (K::NIN >= 2 && (bcast & 4u)) sb = b[0];
  if (K::NIN >= 3 && (bcast & 8u)) sc = c[0];
  if (tile0 + TILE <= n) {
    S1 va[UNROLL][IPT]; S2 vb[UNROLL][IPT]; S3 vc[UNROLL][IPT];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const int64_t e = tile0 + ((int64_t)u * NXC_MAP_THREADS + threadIdx.x) * IPT;
      if (K::NIN >= 1) {
        if (bcast & 2u) { _Pragma("unroll") for (int i = 0; i < IPT; i++) va[u][i] = sa; }
        else nxc_load_vec<S1, IPT>(a + e, va[u]);
      }
      if (K::NIN >= 2) {
        if (bcast & 4u) { _Pragma("unroll") for (int i = 0; i < IPT; i++) vb[u][i] = sb; }
        else nxc_load_vec<S2, IPT>(b + e, vb[u]);
      }
      if (K::NIN >= 3) {
        if (bcast & 8u) { _Pragma("unroll") for (int i = 0; i < IPT; i++) vc[u][i] = sc; }
        else nxc_load_vec<S3, IPT>(c + e, vc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const int64_t e = tile0 + ((int64_t)u * NXC_MAP_THREADS + threadIdx.x) * IPT;
      S0 vo[IPT];
#pragma unroll
      for (int i = 0; i < IPT; i++) vo[i] = K::run(va[u][i], vb[u][i], vc[u][i], prm);
      nxc_store_vec<S0, IPT>(out + e, vo);
    }
  } else {
    for (int64_t e = tile0 + threadIdx.x; e < n; e += NXC_MAP_THREADS) {
      S1 x = sa; S2 y = sb; S3 z = sc;
      if (K::NIN >= 1 && !(bcast & 2u)) x = a[e];
      if (K::NIN >= 2 && !(bcast & 4u)) y = b[e];
      if (K::NIN >= 3 && !(bcast & 8u)) z = c[e];
      out[e] = K::run(x, y, z, prm);
    }
  }
}

// ---- strided kernel ------------------------------------------------------------------
template <int NOP>
__device__ __forceinline__ void nxc_strided_offsets(const NxcStridedArgs<NOP> &args, int64_t item, int64_t (&off)[NOP]) {
  if (args.nitems < 0x7FFFFFFFLL) {
    uint32_t r = (uint32_t)item;
    uint32_t q = nxc_fastdiv(r, args.inner_div);
    uint32_t in = r - q * args.inner_div.d;
#pragma unroll
    for (int k = 0; k < NOP; k++) off[k] = (int64_t)in * args.inner_stride[k];
    r = q;
    for (int d = args.ndim - 1; d >= 0; d--) {
      q = nxc_fastdiv(r, args.div[d]);
      uint32_t cd = r - q * args.div[d].d;
#pragma unroll
      for (int k = 0; k < NOP; k++) off[k] += (int64_t)cd * args.stride[k][d];
      r = q;
    }
  } else {
    int64_t r = item;
    int64_t q = r / args.inner64;
    int64_t in = r - q * args.inner64;
#pragma unroll
    for (int k = 0; k < NOP; k++) off[k] = in * args.inner_stride[k];
    r = q;
    for (int d = args.ndim - 1; d >= 0; d--) {
      q = r / args.shape64[d];
      int64_t cd = r - q * args.shape64[d];
#pragma unroll
      for (int k = 0; k < NOP; k++) off[k] += cd * args.stride[k][d];
      r = q;
    }
  }
}

// U independent work items per thread per iteration: all loads are issued before any
// result is computed, so each thread keeps U x (operands) requests in flight.
template <class K, int VW>
__global__ void __launch_bounds__(NXC_MAP_THREADS)
nxc_map_strided_kernel(typename K::S0 *__restrict__ out, const typename K::S1 *__restrict__ a,
                       const typename K::S2 *__restrict__ b, const typename K::S3 *__restrict__ c,
                       const __grid_constant__ NxcStridedArgs<K::NIN + 1> args, typename K::P prm) {
  typedef typename K::S0 S0; typedef typename K::S1 S1; typedef typename K::S2 S2; typedef typename K::S3 S3;
  constexpr int NOP = K::NIN + 1;
  constexpr int U = (VW * NxcKInfo<K>::max_size >= 64) ? 1 : ((VW * NxcKInfo<K>::max_size >= 16) ? 2 : 4);
  const int64_t step = (int64_t)gridDim.x * NXC_MAP_THREADS;
  for (int64_t item0 = (int64_t)blockIdx.x * NXC_MAP_THREADS + threadIdx.x; item0 < args.nitems;
       item0 += step * U) {
    S1 va[U][VW]; S2 vb[U][VW]; S3 vc[U][VW];
    int64_t oo[U];
    bool live[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t item = item0 + u * step;
      oo[u] = 0;
      live[u] = item < args.nitems;
      if (live[u]) {
        int64_t off[NOP];
        nxc_strided_offsets<NOP>(args, item, off);
        oo[u] = off[0];
        if (K::NIN >= 1) {
          const int64_t o = off[1 < NOP ? 1 : 0];
          if (VW == 1 || (args.bcast_mask & 2u)) { S1 s = a[o]; _Pragma("unroll") for (int i = 0; i < VW; i++) va[u][i] = s; }
          else nxc_load_vec<S1, VW>(a + o, va[u]);
        }
        if (K::NIN >= 2) {
          const int64_t o = off[2 < NOP ? 2 : 0];
          if (VW == 1 || (args.bcast_mask & 4u)) { S2 s = b[o]; _Pragma("unroll") for (int i = 0; i < VW; i++) vb[u][i] = s; }
          else nxc_load_vec<S2, VW>(b + o, vb[u]);
        }
        if (K::NIN >= 3) {
          const int64_t o = off[3 < NOP ? 3 : 0];
          if (VW == 1 || (args.bcast_mask & 8u)) { S3 s = c[o]; _Pragma("unroll") for (int i = 0; i < VW; i++) vc[u][i] = s; }
          else nxc_load_vec<S3, VW>(c + o, vc[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (live[u]) {
        S0 vo[VW];
#pragma unroll
        for (int i = 0; i < VW; i++) vo[i] = K::run(va[u][i], vb[u][i], vc[u][i], prm);
        if (VW == 1) out[oo[u]] = vo[0];
        else nxc_store_vec<S0, VW>(out + oo[u], vo);
      }
    }
  }
}

// ---- tiled kernel: an operand whose unit stride lies on another dim than the output's ---
// (a transposed view). 32x32 tiles over (dim J, inner dim I); operands that are J-major are
// read J-fastest (coalesced), parked in padded shared memory and consumed I-fastest, so both
// the transposed loads and the output stores are full 128-byte warp requests.
template <int NOP>
struct NxcTiledArgs {
  int nrest;                       // dims other than J and I
  NxcFastDiv rest_div[NXC_MAX_NDIM];
  int64_t rest_stride[NOP][NXC_MAX_NDIM];
  int64_t sj[NOP], si[NOP];        // element strides along J and I
  uint32_t SJ, SI, tiles_j, tiles_i;
  NxcFastDiv tiles_i_div, tiles_ij_div;
  uint32_t tmask;                  // bit k: operand k goes through shared memory
};

template <class K>
__global__ void __launch_bounds__(256)
nxc_map_tiled_kernel(typename K::S0 *__restrict__ out, const typename K::S1 *__restrict__ a,
                     const typename K::S2 *__restrict__ b, const typename K::S3 *__restrict__ c,
                     const __grid_constant__ NxcTiledArgs<K::NIN + 1> g, typename K::P prm) {
  typedef typename K::S0 S0; typedef typename K::S1 S1; typedef typename K::S2 S2; typedef typename K::S3 S3;
  constexpr int NOP = K::NIN + 1;
  __shared__ S1 ta[K::NIN >= 1 ? 32 : 1][33];
  __shared__ S2 tb[K::NIN >= 2 ? 32 : 1][33];
  __shared__ S3 tc[K::NIN >= 3 ? 32 : 1][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  uint32_t t = blockIdx.x;
  const uint32_t rest = nxc_fastdiv(t, g.tiles_ij_div);
  t -= rest * g.tiles_ij_div.d;
  const uint32_t tj = nxc_fastdiv(t, g.tiles_i_div);
  const uint32_t ti = t - tj * g.tiles_i_div.d;
  int64_t base[NOP];
#pragma unroll
  for (int k = 0; k < NOP; k++) base[k] = 0;
  {
    uint32_t r = rest;
    for (int d = g.nrest - 1; d >= 0; d--) {
      uint32_t q = nxc_fastdiv(r, g.rest_div[d]);
      uint32_t cd = r - q * g.rest_div[d].d;
#pragma unroll
      for (int k = 0; k < NOP; k++) base[k] += (int64_t)cd * g.rest_stride[k][d];
      r = q;
    }
  }
  const uint32_t j0 = tj * 32, i0 = ti * 32;
  // phase 1: J-major operands -> shared memory (thread x runs along J)
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint32_t il = ty + 8 * r, gi = i0 + il, gj = j0 + tx;
    if (gi < g.SI && gj < g.SJ) {
      if (K::NIN >= 1 && (g.tmask & 2u)) ta[K::NIN >= 1 ? il : 0][tx] = a[base[1 < NOP ? 1 : 0] + (int64_t)gj * g.sj[1 < NOP ? 1 : 0] + (int64_t)gi * g.si[1 < NOP ? 1 : 0]];
      if (K::NIN >= 2 && (g.tmask & 4u)) tb[K::NIN >= 2 ? il : 0][tx] = b[base[2 < NOP ? 2 : 0] + (int64_t)gj * g.sj[2 < NOP ? 2 : 0] + (int64_t)gi * g.si[2 < NOP ? 2 : 0]];
      if (K::NIN >= 3 && (g.tmask & 8u)) tc[K::NIN >= 3 ? il : 0][tx] = c[base[3 < NOP ? 3 : 0] + (int64_t)gj * g.sj[3 < NOP ? 3 : 0] + (int64_t)gi * g.si[3 < NOP ? 3 : 0]];
    }
  }
  __syncthreads();
  // phase 2: thread x runs along I (the output's unit-stride dim)
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint32_t jl = ty + 8 * r, gj = j0 + jl, gi = i0 + tx;
    if (gi < g.SI && gj < g.SJ) {
      S1 va = S1(); S2 vb = S2(); S3 vc = S3();
      if (K::NIN >= 1) va = (g.tmask & 2u) ? ta[K::NIN >= 1 ? tx : 0][jl] : a[base[1 < NOP ? 1 : 0] + (int64_t)gj * g.sj[1 < NOP ? 1 : 0] + (int64_t)gi * g.si[1 < NOP ? 1 : 0]];
      if (K::NIN >= 2) vb = (g.tmask & 4u) ? tb[K::NIN >= 2 ? tx : 0][jl] : b[base[2 < NOP ? 2 : 0] + (int64_t)gj * g.sj[2 < NOP ? 2 : 0] + (int64_t)gi * g.si[2 < NOP ? 2 : 0]];
      if (K::NIN >= 3) vc = (g.tmask & 8u) ? tc[K::NIN >= 3 ? tx : 0][jl] : c[base[3 < NOP ? 3 : 0] + (int64_t)gj * g.sj[3 < NOP ? 3 : 0] + (int64_t)gi * g.si[3 < NOP ? 3 : 0]];
      out[base[0] + (int64_t)gj * g.sj[0] + (int64_t)gi * g.si[0]] = K::run(va, vb, vc, prm);
    }
  }
}

// Host side: does this plan want the tiled kernel? Picks J = the dim on which some input has
// unit stride while the output's unit stride is on the last dim.
template <class K, bool ENABLED> struct NxcTiledLaunch {
  static bool go(nxc_ctx *, const NxcMapPlan &, typename K::P, nxc_status *) { return false; }
};
template <class K> struct NxcTiledLaunch<K, true> {
  static bool go(nxc_ctx *ctx, const NxcMapPlan &p, typename K::P prm, nxc_status *st) {
    typedef typename K::S0 S0; typedef typename K::S1 S1; typedef typename K::S2 S2; typedef typename K::S3 S3;
    constexpr int NOP = K::NIN + 1;
    const int I = p.ndim - 1;
    if (p.ndim < 2 || p.stride[0][I] != 1 || p.shape[I] < 16 || p.total >= 0x7FFFFFFFLL * 64) return false;
    int J = -1;
    uint32_t tmask = 0;
    for (int k = 1; k < NOP; k++) {
      if (p.stride[k][I] == 1 || p.stride[k][I] == 0) continue;
      for (int d = 0; d < I; d++)
        if (p.stride[k][d] == 1 && p.shape[d] >= 16 && (J < 0 || J == d)) { J = d; tmask |= 1u << k; }
    }
    if (J < 0) return false;
    NxcTiledArgs<NOP> g;
    g.SJ = (uint32_t)p.shape[J];
    g.SI = (uint32_t)p.shape[I];
    if (p.shape[J] > 0x7FFFFFFF || p.shape[I] > 0x7FFFFFFF) return false;
    g.tiles_j = (g.SJ + 31) / 32;
    g.tiles_i = (g.SI + 31) / 32;
    g.tmask = tmask;
    int64_t nrest_total = 1;
    g.nrest = 0;
    for (int d = 0; d < I; d++) {
      if (d == J) continue;
      g.rest_div[g.nrest] = nxc_fastdiv_make((uint32_t)p.shape[d]);
      for (int k = 0; k < NOP; k++) g.rest_stride[k][g.nrest] = p.stride[k][d];
      nrest_total *= p.shape[d];
      g.nrest++;
    }
    for (int k = 0; k < NOP; k++) { g.sj[k] = p.stride[k][J]; g.si[k] = p.stride[k][I]; }
    const int64_t blocks = nrest_total * g.tiles_j * g.tiles_i;
    if (blocks >= 0x7FFFFFFFLL) return false;
    g.tiles_i_div = nxc_fastdiv_make(g.tiles_i);
    g.tiles_ij_div = nxc_fastdiv_make(g.tiles_i * g.tiles_j);
    nxc_map_tiled_kernel<K><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
        (S0 *)p.base[0], (const S1 *)(NOP > 1 ? p.base[1] : p.base[0]), (const S2 *)(NOP > 2 ? p.base[2] : p.base[0]),
        (const S3 *)(NOP > 3 ? p.base[3] : p.base[0]), g, prm);
    ctx->launches++;
    cudaError_t e = cudaPeekAtLastError();
    *st = (e == cudaSuccess) ? NXC_OK : nxc_cuda_fail(ctx, e, "kernel launch");
    return true;
  }
};
template <class K, class = void> struct NxcIsTiled { static constexpr bool v = false; };
template <class K> struct NxcIsTiled<K, typename std::enable_if<K::TILED>::type> { static constexpr bool v = true; };

// ---- launcher ------------------------------------------------------------------------
static inline bool nxc_aligned(const void *p, size_t a) { return ((uintptr_t)p % a) == 0; }

template <class K>
nxc_status nxc_map_launch(nxc_ctx *ctx, const NxcMapPlan &p, typename K::P prm) {
  typedef typename K::S0 S0; typedef typename K::S1 S1; typedef typename K::S2 S2; typedef typename K::S3 S3;
  constexpr int NOP = K::NIN + 1;
  constexpr int IPT = NxcKInfo<K>::IPT, UNROLL = NxcKInfo<K>::UNROLL;
  if (p.total == 0) return NXC_OK;
  S0 *o = (S0 *)p.base[0];
  const S1 *a = (const S1 *)(NOP > 1 ? p.base[1] : p.base[0]);
  const S2 *b = (const S2 *)(NOP > 2 ? p.base[2] : p.base[0]);
  const S3 *c = (const S3 *)(NOP > 3 ? p.base[3] : p.base[0]);
  const size_t esz[4] = {sizeof(S0), sizeof(S1), sizeof(S2), sizeof(S3)};

  // flat path?
  bool flat = (p.ndim == 1) && p.stride[0][0] == 1;
  uint32_t bc = 0;
  if (flat) {
    for (int k = 0; k < NOP; k++) {
      int64_t s = p.stride[k][0];
      if (s == 0 && k > 0) bc |= 1u << k;
      else if (s != 1) flat = false;
      else if (!nxc_aligned(p.base[k], esz[k] * IPT)) flat = false;
    }
  }
  if (p.total == 1) { flat = true; bc = 0; }  // single element: tail loop handles it
  if (flat) {
    const int64_t tile = (int64_t)NXC_MAP_THREADS * IPT * UNROLL;
    const int64_t blocks = (p.total + tile - 1) / tile;
    nxc_map_flat_kernel<K><<<(unsigned)blocks, NXC_MAP_THREADS, 0, ctx->stream>>>(o, a, b, c, p.total, bc, prm);
    NXC_LAUNCH_CHECK(ctx);
    return NXC_OK;
  }

  // transposed operand -> shared-memory tiles (ops that opt in)
  {
    nxc_status tst = NXC_OK;
    if (NxcTiledLaunch<K, NxcIsTiled<K>::v>::go(ctx, p, prm, &tst)) return tst;
  }

  // strided path
  NxcStridedArgs<NOP> g;
  const int od = p.ndim - 1;
  int64_t inner = p.shape[od];
  // vector width over the inner dim
  int vw = 1;
  if (IPT > 1 && inner % IPT == 0) {
    bool ok = true;
    for (int k = 0; k < NOP && ok; k++) {
      int64_t s = p.stride[k][od];
      if (k == 0 ? s != 1 : (s != 0 && s != 1)) ok = false;
      if (s == 1) {
        if (!nxc_aligned(p.base[k], esz[k] * IPT)) ok = false;
        for (int d = 0; d < od && ok; d++)
          if (p.stride[k][d] % IPT != 0) ok = false;
      }
    }
    if (ok) vw = IPT;
  }
  g.ndim = od;
  g.bcast_mask = 0;
  const int64_t inner_items = inner / vw;
  g.inner = (uint32_t)inner_items;
  g.inner64 = inner_items;
  for (int k = 0; k < NOP; k++) {
    g.inner_stride[k] = p.stride[k][od] * vw;
    if (p.stride[k][od] == 0) g.bcast_mask |= 1u << k;
    for (int d = 0; d < od; d++) g.stride[k][d] = p.stride[k][d];
  }
  g.nitems = p.total / vw;
  const bool small = g.nitems < 0x7FFFFFFFLL;
  g.inner_div = nxc_fastdiv_make(small ? (uint32_t)inner_items : 1u);
  for (int d = 0; d < od; d++) {
    g.shape64[d] = p.shape[d];
    g.div[d] = nxc_fastdiv_make(small ? (uint32_t)p.shape[d] : 1u);
  }
  int64_t blocks = (g.nitems + NXC_MAP_THREADS - 1) / NXC_MAP_THREADS;
  const int64_t cap = (int64_t)ctx->sm_count * 32;
  if (blocks > cap) blocks = cap;
  if (vw == 1)
    nxc_map_strided_kernel<K, 1><<<(unsigned)blocks, NXC_MAP_THREADS, 0, ctx->stream>>>(o, a, b, c, g, prm);
  else
    nxc_map_strided_kernel<K, IPT><<<(unsigned)blocks, NXC_MAP_THREADS, 0, ctx->stream>>>(o, a, b, c, g, prm);
  NXC_LAUNCH_CHECK(ctx);
  return NXC_OK;
}

// Launch only when the (op, dtype) slot exists; the discarded side is never
// instantiated (a plain `if constexpr` in a non-template would still compile it).
template <class K, bool OK> struct NxcMaybeMap {
  static nxc_status go(nxc_ctx *ctx, const NxcMapPlan &p, typename K::P prm) { return nxc_map_launch<K>(ctx, p, prm); }
};
template <class K> struct NxcMaybeMap<K, false> {
  static nxc_status go(nxc_ctx *, const NxcMapPlan &, typename K::P) { return NXC_ERR_UNSUPPORTED_DTYPE; }
};
