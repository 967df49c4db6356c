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
template <class K, int VW>
__global__ void __launch_bounds__(NXC_MAP_THREADS)
nxc_map_strided_kernel(typename K::S0 *__restrict__ out, const typename K::S1 *__restrict__ a,
                       const typename K::S2 *__restrict__ b, const typename K::S3 *__restrict__ c,
                       const __grid_constant__ NxcStridedArgs<K::NIN + 1> args, typename K::P prm) {
  typedef typename K::S0 S0; typedef typename K::S1 S1; typedef typename K::S2 S2; typedef typename K::S3 S3;
  constexpr int NOP = K::NIN + 1;
  const int64_t step = (int64_t)gridDim.x * NXC_MAP_THREADS;
  for (int64_t item = (int64_t)blockIdx.x * NXC_MAP_THREADS + threadIdx.x; item < args.nitems;
       item += step) {
    int64_t off[NOP];
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
    S1 va[VW]; S2 vb[VW]; S3 vc[VW]; S0 vo[VW];
    if (VW == 1) {
      if (K::NIN >= 1) va[0] = a[off[1 < NOP ? 1 : 0]];
      if (K::NIN >= 2) vb[0] = b[off[2 < NOP ? 2 : 0]];
      if (K::NIN >= 3) vc[0] = c[off[3 < NOP ? 3 : 0]];
      out[off[0]] = K::run(va[0], vb[0], vc[0], prm);
    } else {
      if (K::NIN >= 1) {
        const int64_t o = off[1 < NOP ? 1 : 0];
        if (args.bcast_mask & 2u) { S1 s = a[o]; _Pragma("unroll") for (int i = 0; i < VW; i++) va[i] = s; }
        else nxc_load_vec<S1, VW>(a + o, va);
      }
      if (K::NIN >= 2) {
        const int64_t o = off[2 < NOP ? 2 : 0];
        if (args.bcast_mask & 4u) { S2 s = b[o]; _Pragma("unroll") for (int i = 0; i < VW; i++) vb[i] = s; }
        else nxc_load_vec<S2, VW>(b + o, vb);
      }
      if (K::NIN >= 3) {
        const int64_t o = off[3 < NOP ? 3 : 0];
        if (args.bcast_mask & 8u) { S3 s = c[o]; _Pragma("unroll") for (int i = 0; i < VW; i++) vc[i] = s; }
        else nxc_load_vec<S3, VW>(c + o, vc);
      }
#pragma unroll
      for (int i = 0; i < VW; i++) vo[i] = K::run(va[i], vb[i], vc[i], prm);
      nxc_store_vec<S0, VW>(out + off[0], vo);
    }
  }
}

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
