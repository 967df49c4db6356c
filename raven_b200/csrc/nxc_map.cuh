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
  int ndim;                 // outer dims ("rows"); the inner dim is handled separately
  uint32_t bcast_mask;      // bit k: operand k has inner stride 0
  uint32_t neg_mask;        // bit k: operand k runs backwards along the inner dim (vector mode)
  int tx_log2;              // threads along the inner dim; 256 >> tx_log2 rows per CTA pass
  int ui_log2;              // of the 4 items a thread owns, 2^ui_log2 lie along the inner dim
  uint32_t chunks;          // CTAs along the inner dim
  NxcFastDiv chunks_div;
  NxcFastDiv div[NXC_MAX_NDIM];
  int64_t shape64[NXC_MAX_NDIM];
  int64_t stride[NOP][NXC_MAX_NDIM];  // outer strides, elements
  int64_t inner_stride[NOP];          // elements per work item
  int64_t rows;                       // product of the outer extents
  int64_t ni;                         // inner extent in work items
};

static inline bool nxc_aligned(const void *p, size_t a) { return ((uintptr_t)p % a) == 0; }

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
  // items per thread per step: the smallest type moves 16 bytes, but the largest no more than
  // 64 (a thread's 16-byte loads of one operand then stay within two 32-byte sectors' reach of
  // its neighbours'; 8-byte -> bool at 16 items per thread measured 0.71 of the f32 rate)
  static constexpr int IPT_RAW = 16 / min_size;
  // A cast whose sides differ 4x or more in width (i8 <-> f32, i16 <-> f64 ...): let the WIDE side move
  // 16 bytes per lane -- the warp's wide accesses are then one contiguous 512 bytes and the narrow ones
  // one contiguous 128 -- instead of the narrow side (one 16-byte load feeding four 16-byte stores 64
  // bytes apart from lane to lane: every store instruction touched 32 quarter-filled segments, 0.70).
  static constexpr bool WIDE_SIDE = K::NIN == 1 && max_size >= 4 * min_size;
  static constexpr int IPT = WIDE_SIDE ? 16 / max_size : (IPT_RAW * max_size > 64 ? 64 / max_size : IPT_RAW);
  // independent steps in flight: aim at 64 bytes of the largest type per thread
  // ... and at least 32 bytes of the SMALLEST: a widening cast (i8 -> f32: one 16-byte load feeding
  // four 16-byte stores) otherwise keeps a single load in flight per thread (measured 0.70)
  static constexpr int UNROLL_BIG = 64 / (IPT * max_size), UNROLL_SMALL = 32 / (IPT * min_size);
  static constexpr int UNROLL_RAW = UNROLL_BIG > UNROLL_SMALL ? UNROLL_BIG : UNROLL_SMALL;
  static constexpr int UNROLL = UNROLL_RAW < 1 ? 1 : (UNROLL_RAW > 4 ? 4 : UNROLL_RAW);
};

#define NXC_MAP_THREADS 256

// N results of one op. Ops that declare IN_DT / OUT_DT / op() (unary, binary, casts to a float type)
// and touch a 16-bit float type convert in bulk: f16 inputs through nxc_ld_many, f16 / bf16 results
// through nxc_pack16 (two per hardware convert); everything else is N scalar calls.
template <class K, class = void> struct NxcBulk { static constexpr bool v = false; };
template <class K> struct NxcBulk<K, typename std::enable_if<(K::OUT_DT >= 0)>::type> {
  static constexpr bool v = K::OUT_DT == NXC_F16 || K::OUT_DT == NXC_BF16 || K::IN_DT == NXC_F16;
};
template <class K, int N>
__device__ __forceinline__ void nxc_run_vec(const typename K::S1 (&a)[N], const typename K::S2 (&b)[N],
                                            const typename K::S3 (&c)[N], typename K::S0 (&o)[N], const typename K::P &prm) {
  if constexpr (N % 2 == 0 && NxcBulk<K>::v) {
    constexpr int IDT = K::IN_DT, ODT = K::OUT_DT;
    typedef typename DT_<IDT>::C CI;
    typedef typename DT_<ODT>::C CO;
    CI ca[N], cb[N];
    nxc_ld_many<IDT, N>(a, ca);
    if constexpr (K::NIN >= 2) nxc_ld_many<IDT, N>(b, cb);
    CO v[N];
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = K::op(ca[i], K::NIN >= 2 ? cb[i] : ca[i], prm);
    if constexpr (ODT == NXC_F16 || ODT == NXC_BF16) {
      nxc_pack16<ODT, N>(v, o);
    } else {
#pragma unroll
      for (int i = 0; i < N; i++) o[i] = DT_<ODT>::st(v[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; i++) o[i] = K::run(a[i], b[i], c[i], prm);
  }
}

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
      nxc_run_vec<K, IPT>(va[u], vb[u], vc[u], vo, prm);
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
// The coalesced plan is seen as rows x inner: a CTA covers (256/TX * UR) rows by (TX * UI) inner
// work items, UI * UR = 4, a work item = VW consecutive inner elements. A thread decodes its row
// coordinates once (32-bit multiply-shift division per outer dim) and then only adds multiples
// of the inner stride, so there is no per-element division; all 4 items' loads are issued
// before any result is computed.
// one operand's work item: VW elements (forwards, or the VW elements ending at p for a reversed
// run) or one broadcast element in r[0]. Nothing here CONSUMES a loaded value: a register move
// after the load would make the thread wait for that load before issuing the next slot's (ncu:
// 72 % of the stall samples sat on the broadcast moves) -- nxc_item_elem picks the element at
// compute time instead.
template <typename S, int VW>
__device__ __forceinline__ void nxc_load_item(const S *p, bool bcast, bool neg, S (&r)[VW], S &sc) {
  // A broadcast element gets a register of its OWN: sharing r[0] between the scalar and the vector
  // load makes ptxas merge the two predicated destinations with a move placed right behind the load
  // -- for the 8-byte types that put one dependent round trip per slot back (ncu, f64
  // [R,C]+[R,1]: four moves with 20 % of the stall samples each, 0.77 of the roofline).
  if (VW == 1) r[0] = *p;
  else if (bcast) sc = *p;
  else nxc_load_vec<S, VW>(neg ? p - (VW - 1) : p, r);
}
template <typename S, int VW>
__device__ __forceinline__ S nxc_item_elem(const S (&r)[VW], const S &sc, int i, bool bcast, bool neg) {
  if (VW == 1) return r[0];
  return bcast ? sc : (neg ? r[VW - 1 - i] : r[i]);
}

// row -> element offsets of the first NEED operands (the outer coordinates' contribution)
template <int NOP, int NEED>
__device__ __forceinline__ void nxc_row_offsets(const NxcStridedArgs<NOP> &args, int64_t row, int64_t (&off)[NEED]) {
  if (args.ndim == 1) {
#pragma unroll
    for (int k = 0; k < NEED; k++) off[k] = row * args.stride[k][0];
    return;
  }
#pragma unroll
  for (int k = 0; k < NEED; k++) off[k] = 0;
  if (args.rows < 0x7FFFFFFFLL) {
    uint32_t r = (uint32_t)row;
    for (int d = args.ndim - 1; d >= 0; d--) {
      const uint32_t q = nxc_fastdiv(r, args.div[d]);
      const uint32_t cd = r - q * args.div[d].d;
#pragma unroll
      for (int k = 0; k < NEED; k++) off[k] += (int64_t)cd * args.stride[k][d];
      r = q;
    }
  } else {
    int64_t r = row;
    for (int d = args.ndim - 1; d >= 0; d--) {
      const int64_t q = r / args.shape64[d];
      const int64_t cd = r - q * args.shape64[d];
#pragma unroll
      for (int k = 0; k < NEED; k++) off[k] += cd * args.stride[k][d];
      r = q;
    }
  }
}

// ROWS = false: long inner dim. A thread owns ONE row (coordinates decoded once) and 4 work
// items TX apart along it. ROWS = true: short inner dim. A thread owns one inner position of 4
// rows TY apart; the output offset is recomputed at store time instead of being kept alive
// across the loads (registers decide how many CTAs fit on an SM, and that decides bandwidth).
template <class K, int VW, bool ROWS>
__global__ void __launch_bounds__(NXC_MAP_THREADS)
nxc_map_strided_kernel(typename K::S0 *__restrict__ out, const typename K::S1 *__restrict__ a,
                       const typename K::S2 *__restrict__ b, const typename K::S3 *__restrict__ c,
                       const __grid_constant__ NxcStridedArgs<K::NIN + 1> args, typename K::P prm) {
  typedef typename K::S0 S0; typedef typename K::S1 S1; typedef typename K::S2 S2; typedef typename K::S3 S3;
  constexpr int NOP = K::NIN + 1;
  constexpr int U = 4;
  constexpr int KA = 1 < NOP ? 1 : 0, KB = 2 < NOP ? 2 : 0, KC = 3 < NOP ? 3 : 0;
  const uint32_t tx = threadIdx.x & ((1u << args.tx_log2) - 1u), ty = threadIdx.x >> args.tx_log2;
  const uint32_t TX = 1u << args.tx_log2, TY = NXC_MAP_THREADS >> args.tx_log2;
  const uint32_t rb = nxc_fastdiv(blockIdx.x, args.chunks_div);
  const uint32_t chunk = blockIdx.x - rb * args.chunks;
  S1 va[U][VW]; S2 vb[U][VW]; S3 vc[U][VW];
  S1 xa[U]; S2 xb[U]; S3 xc[U];  // broadcast elements (see nxc_load_item)
  const bool ba = args.bcast_mask & 2u, bb = args.bcast_mask & 4u, bc = args.bcast_mask & 8u;
  const bool na = args.neg_mask & 2u, nb = args.neg_mask & 4u, nc = args.neg_mask & 8u;
  // the 4 slots of a thread are a constant pointer step apart (TX items along the row, or TY
  // rows): pointers advance by addition, and "slot u exists" is u < n_live
  S0 *po; const S1 *pa; const S2 *pb; const S3 *pc;
  int64_t so, sa, sb, sc;
  int n_live;
  if (!ROWS) {
    const int64_t row = (int64_t)rb * TY + ty;
    const int64_t in0 = (int64_t)chunk * (TX * U) + tx;
    if (row >= args.rows || in0 >= args.ni) return;
    int64_t off[NOP];
    nxc_row_offsets<NOP, NOP>(args, row, off);
    po = out + off[0] + in0 * args.inner_stride[0];
    pa = a + off[KA] + in0 * args.inner_stride[KA];
    pb = b + off[KB] + in0 * args.inner_stride[KB];
    pc = c + off[KC] + in0 * args.inner_stride[KC];
    so = (int64_t)TX * args.inner_stride[0]; sa = (int64_t)TX * args.inner_stride[KA];
    sb = (int64_t)TX * args.inner_stride[KB]; sc = (int64_t)TX * args.inner_stride[KC];
    const int64_t left = (args.ni - in0 + TX - 1) >> args.tx_log2;
    n_live = left < U ? (int)left : U;
  } else if (args.ndim <= 1) {
    const int64_t item = (int64_t)chunk * TX + tx;
    const int64_t row0 = (int64_t)rb * (TY * U) + ty;
    if (item >= args.ni || row0 >= args.rows) return;
    const int64_t r0 = args.ndim ? args.stride[0][0] : 0, ra = args.ndim ? args.stride[KA][0] : 0,
                  rb_ = args.ndim ? args.stride[KB][0] : 0, rc = args.ndim ? args.stride[KC][0] : 0;
    po = out + row0 * r0 + item * args.inner_stride[0];
    pa = a + row0 * ra + item * args.inner_stride[KA];
    pb = b + row0 * rb_ + item * args.inner_stride[KB];
    pc = c + row0 * rc + item * args.inner_stride[KC];
    so = (int64_t)TY * r0; sa = (int64_t)TY * ra; sb = (int64_t)TY * rb_; sc = (int64_t)TY * rc;
    const int64_t left = (args.rows - row0 + TY - 1) / TY;
    n_live = left < U ? (int)left : U;
  } else {
    n_live = -1;
  }
  if (!ROWS || n_live >= 0) {
    // Load phase: nothing predicated. Slots past the end re-load the last live slot (their results
    // are never stored) and the broadcast pattern, uniform over the grid, picks one of a few
    // straight-line variants by a real branch. With `@p LDG` forms ptxas funnels the 8-byte types'
    // vector loads through one register quad and copies each result out right behind its load --
    // one dependent round trip per slot (ncu, f64 [R,C]+[R,1]: 0.77 of the roofline, four moves
    // holding 20 % of the stall samples each).
    // The same flags are compile-time constants in the COMPUTE phase of those variants: picking an
    // element out of a packed register by a run-time "broadcast? reversed?" costs two selects per
    // operand and element, which for the 1- and 2-byte types (8-16 elements per 16-byte item) made the
    // kernel ALU-bound (bf16 [R,C]+[1,C]: 0.49 of the roofline).
    auto run_all = [&](auto BA, auto BB, auto BC, auto NG) {
      constexpr bool cba = decltype(BA)::value, cbb = decltype(BB)::value, cbc = decltype(BC)::value;
      constexpr bool ng = decltype(NG)::value;
      const bool na_ = ng && na, nb_ = ng && nb, nc_ = ng && nc;
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (K::NIN >= 1) nxc_load_item<S1, VW>(pa, cba, na_, va[u], xa[u]);
        if (K::NIN >= 2) nxc_load_item<S2, VW>(pb, cbb, nb_, vb[u], xb[u]);
        if (K::NIN >= 3) nxc_load_item<S3, VW>(pc, cbc, nc_, vc[u], xc[u]);
        if (u + 1 < n_live) { pa += sa; pb += sb; pc += sc; }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (u < n_live) {
          S0 vo[VW];
          S1 ea[VW]; S2 eb[VW]; S3 ec[VW];
#pragma unroll
          for (int i = 0; i < VW; i++) {
            ea[i] = nxc_item_elem<S1, VW>(va[u], xa[u], i, cba, na_);
            eb[i] = nxc_item_elem<S2, VW>(vb[u], xb[u], i, cbb, nb_);
            ec[i] = nxc_item_elem<S3, VW>(vc[u], xc[u], i, cbc, nc_);
          }
          nxc_run_vec<K, VW>(ea, eb, ec, vo, prm);
          if (VW == 1) po[0] = vo[0];
          else nxc_store_vec<S0, VW>(po, vo);
        }
        po += so;
      }
    };
    typedef std::integral_constant<bool, false> F_;
    typedef std::integral_constant<bool, true> T_;
    const bool any_neg = na | nb | nc;
    // a scalar item (VW == 1) is loaded and consumed the same way whatever the flags: one body
    // (the straight-line variants of it were identical code, a third of libnxcuda.so's size)
    if (VW > 1 && !any_neg && !(ba | bb | bc)) run_all(F_(), F_(), F_(), F_());
    else if (VW > 1 && !any_neg && bb && !ba && !bc) run_all(F_(), T_(), F_(), F_());
    else if (VW > 1 && !any_neg && ba && !bb && !bc) run_all(T_(), F_(), F_(), F_());
    else {
      // anything else (a reversed run, several broadcast operands): flags stay run-time values
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (K::NIN >= 1) nxc_load_item<S1, VW>(pa, ba, na, va[u], xa[u]);
        if (K::NIN >= 2) nxc_load_item<S2, VW>(pb, bb, nb, vb[u], xb[u]);
        if (K::NIN >= 3) nxc_load_item<S3, VW>(pc, bc, nc, vc[u], xc[u]);
        if (u + 1 < n_live) { pa += sa; pb += sb; pc += sc; }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (u < n_live) {
          S0 vo[VW];
          S1 ea[VW]; S2 eb[VW]; S3 ec[VW];
#pragma unroll
          for (int i = 0; i < VW; i++) {
            ea[i] = nxc_item_elem<S1, VW>(va[u], xa[u], i, ba, na);
            eb[i] = nxc_item_elem<S2, VW>(vb[u], xb[u], i, bb, nb);
            ec[i] = nxc_item_elem<S3, VW>(vc[u], xc[u], i, bc, nc);
          }
          nxc_run_vec<K, VW>(ea, eb, ec, vo, prm);
          if (VW == 1) po[0] = vo[0];
          else nxc_store_vec<S0, VW>(po, vo);
        }
        po += so;
      }
    }
  } else {
    const int64_t item = (int64_t)chunk * TX + tx;
    if (item >= args.ni) return;
    const int64_t row0 = (int64_t)rb * (TY * U) + ty;
    {
      // several outer dims: coordinates are decoded per row, one row at a time
#pragma unroll 1
      for (int u = 0; u < U; u++) {
        const int64_t row = row0 + (int64_t)(u * TY);
        if (row >= args.rows) break;
        int64_t off[NOP];
        nxc_row_offsets<NOP, NOP>(args, row, off);
        if (K::NIN >= 1) nxc_load_item<S1, VW>(a + off[KA] + item * args.inner_stride[KA], args.bcast_mask & 2u, args.neg_mask & 2u, va[0], xa[0]);
        if (K::NIN >= 2) nxc_load_item<S2, VW>(b + off[KB] + item * args.inner_stride[KB], args.bcast_mask & 4u, args.neg_mask & 4u, vb[0], xb[0]);
        if (K::NIN >= 3) nxc_load_item<S3, VW>(c + off[KC] + item * args.inner_stride[KC], args.bcast_mask & 8u, args.neg_mask & 8u, vc[0], xc[0]);
        S0 vo[VW];
#pragma unroll
        for (int i = 0; i < VW; i++)
          vo[i] = K::run(nxc_item_elem<S1, VW>(va[0], xa[0], i, ba, na), nxc_item_elem<S2, VW>(vb[0], xb[0], i, bb, nb),
                         nxc_item_elem<S3, VW>(vc[0], xc[0], i, bc, nc), prm);
        S0 *q = out + off[0] + item * args.inner_stride[0];
        if (VW == 1) q[0] = vo[0];
        else nxc_store_vec<S0, VW>(q, vo);
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

// Vector variant of the tiled kernel: T x T tiles (64 for elements <= 4 bytes, 32 for 8-byte
// ones), every global access a V-element vector (16 bytes for f32 / f64) -- transposed operands
// are read in 256-byte row segments along THEIR unit-stride dim, scattered into padded shared
// memory and gathered back along the output's; straight operands are loaded before the barrier
// so both streams are in flight together. Needs V-aligned extents, strides and bases; the scalar
// 32x32 kernel above stays as the general fallback.
template <class K> struct NxcTiledV {
  static constexpr int ESZ = (int)sizeof(typename K::S0);
  static constexpr int T = ESZ <= 4 ? 64 : 32;
  static constexpr int V = ESZ >= 8 ? 2 : 4;
  static constexpr int VPR = T / V;     // vectors per tile row
  static constexpr int RP = 256 / VPR;  // tile rows per pass
  static constexpr int NP = T / RP;     // passes
};

template <class K>
__global__ void __launch_bounds__(256)
nxc_map_tiledv_kernel(typename K::S0 *__restrict__ out, const typename K::S1 *__restrict__ a,
                      const typename K::S2 *__restrict__ b, const typename K::S3 *__restrict__ c,
                      const __grid_constant__ NxcTiledArgs<K::NIN + 1> g, typename K::P prm) {
  typedef typename K::S0 S0; typedef typename K::S1 S1; typedef typename K::S2 S2; typedef typename K::S3 S3;
  constexpr int NOP = K::NIN + 1;
  constexpr int T = NxcTiledV<K>::T, V = NxcTiledV<K>::V, VPR = NxcTiledV<K>::VPR, RP = NxcTiledV<K>::RP,
                NP = NxcTiledV<K>::NP;
  static_assert(K::NIN <= 2, "tiled ops have at most two inputs");
  __shared__ S1 ta[K::NIN >= 1 ? T : 1][T + 1];
  __shared__ S2 tb[K::NIN >= 2 ? T : 1][T + 1];
  const int vx = threadIdx.x % VPR, ry = threadIdx.x / VPR;
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
  const uint32_t j0 = tj * T, i0 = ti * T;
  constexpr int KA = 1 < NOP ? 1 : 0, KB = 2 < NOP ? 2 : 0;
  const bool a_t = K::NIN >= 1 && (g.tmask & 2u), b_t = K::NIN >= 2 && (g.tmask & 4u);
  S1 ra[NP][V]; S2 rb[NP][V];
  // transposed operands: vectors along J at fixed i
#pragma unroll
  for (int r = 0; r < NP; r++) {
    const uint32_t il = ry + RP * r, gi = i0 + il, gj = j0 + vx * V;
    if (gi < g.SI && gj < g.SJ) {
      if (a_t) nxc_load_vec<S1, V>(a + base[KA] + (int64_t)gj + (int64_t)gi * g.si[KA], ra[r]);
      if (b_t) nxc_load_vec<S2, V>(b + base[KB] + (int64_t)gj + (int64_t)gi * g.si[KB], rb[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < NP; r++) {
    const uint32_t il = ry + RP * r;
    if (i0 + il < g.SI && j0 + vx * V < g.SJ) {
#pragma unroll
      for (int q = 0; q < V; q++) {
        if (a_t) ta[K::NIN >= 1 ? il : 0][vx * V + q] = ra[r][q];
        if (b_t) tb[K::NIN >= 2 ? il : 0][vx * V + q] = rb[r][q];
      }
    }
  }
  // straight operands: vectors along I at fixed j (stride 1) or one broadcast element (stride 0)
#pragma unroll
  for (int r = 0; r < NP; r++) {
    const uint32_t jl = ry + RP * r, gj = j0 + jl, gi = i0 + vx * V;
    if (gi < g.SI && gj < g.SJ) {
      if (K::NIN >= 1 && !a_t) {
        const S1 *p = a + base[KA] + (int64_t)gj * g.sj[KA];
        if (g.si[KA] == 0) { S1 s = p[0]; _Pragma("unroll") for (int q = 0; q < V; q++) ra[r][q] = s; }
        else nxc_load_vec<S1, V>(p + gi, ra[r]);
      }
      if (K::NIN >= 2 && !b_t) {
        const S2 *p = b + base[KB] + (int64_t)gj * g.sj[KB];
        if (g.si[KB] == 0) { S2 s = p[0]; _Pragma("unroll") for (int q = 0; q < V; q++) rb[r][q] = s; }
        else nxc_load_vec<S2, V>(p + gi, rb[r]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < NP; r++) {
    const uint32_t jl = ry + RP * r, gj = j0 + jl, gi = i0 + vx * V;
    if (gi < g.SI && gj < g.SJ) {
      S0 vo[V];
      S1 ea[V]; S2 eb[V]; S3 ec[V];
#pragma unroll
      for (int q = 0; q < V; q++) {
        ea[q] = S1(); eb[q] = S2(); ec[q] = S3();
        if (K::NIN >= 1) ea[q] = a_t ? ta[K::NIN >= 1 ? vx * V + q : 0][jl] : ra[r][q];
        if (K::NIN >= 2) eb[q] = b_t ? tb[K::NIN >= 2 ? vx * V + q : 0][jl] : rb[r][q];
      }
      nxc_run_vec<K, V>(ea, eb, ec, vo, prm);
      nxc_store_vec<S0, V>(out + base[0] + (int64_t)gj * g.sj[0] + gi, vo);
    }
  }
}


// Narrow variant of the vector tiled kernel, for 1- and 2-byte elements: with those the kernel above
// moves only 4 elements = 4-8 bytes per lane and access (measured 0.44-0.66 of the HBM rate: the LSU
// is the limit, not memory). Here a tile is 128 x 128 elements, EVERY global access is a 16-byte
// vector (8 or 16 elements) and every shared-memory access moves at least 4 bytes:
//   * a transposed operand is stored with 16-byte vector stores, rows unpadded, the 16-byte chunks of
//     a row XOR-swizzled by (row / V), so that the gather in the other direction -- lane vx reads rows
//     vx*V + q, V rows apart -- finds its 16 (8) lanes in 16 (8) different chunks;
//   * the gather reads 32-bit WORDS: a word holds the same source row's elements for W = 2 (4)
//     adjacent output rows, so a thread produces W output vectors at once and issues V word loads
//     for them instead of W*V sub-word ones.
template <class K> struct NxcTiledN {
  static constexpr int ESZ = (int)sizeof(typename K::S0);
  static constexpr int T = 128;
  static constexpr int V = 16 / ESZ;      // elements per 16-byte vector
  static constexpr int W = 4 / ESZ;       // elements per 32-bit word = output rows a thread produces per pass
  static constexpr int CPR = T / V;       // 16-byte chunks (= vector lanes) per tile row
  static constexpr int RP = 256 / CPR;    // thread rows
  static constexpr int NPL = T / RP;      // load passes (one tile row per thread row and pass)
  static constexpr int NPO = T / (RP * W);  // output passes (W tile rows per thread row and pass)
  static constexpr int TILE_BYTES = T * T * ESZ;
};

template <class K>
__global__ void __launch_bounds__(256)
nxc_map_tiledn_kernel(typename K::S0 *__restrict__ out, const typename K::S1 *__restrict__ a,
                      const typename K::S2 *__restrict__ b, const typename K::S3 *__restrict__ c,
                      const __grid_constant__ NxcTiledArgs<K::NIN + 1> g, typename K::P prm) {
  typedef typename K::S0 S0; typedef typename K::S1 S1; typedef typename K::S2 S2; typedef typename K::S3 S3;
  typedef NxcTiledN<K> C;
  constexpr int NOP = K::NIN + 1;
  constexpr int T = C::T, V = C::V, W = C::W, CPR = C::CPR, RP = C::RP, NPL = C::NPL, NPO = C::NPO, ESZ = C::ESZ;
  static_assert(K::NIN >= 1 && K::NIN <= 2, "tiled ops have one or two inputs");
  extern __shared__ __align__(16) unsigned char nxc_tiledn_smem[];
  const int vx = threadIdx.x % CPR, ry = threadIdx.x / CPR;
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
  const uint32_t j0 = tj * T, i0 = ti * T;
  constexpr int KA = 1 < NOP ? 1 : 0, KB = 2 < NOP ? 2 : 0;
  const bool a_t = (g.tmask & 2u) != 0, b_t = K::NIN >= 2 && (g.tmask & 4u) != 0;
  unsigned char *ta = nxc_tiledn_smem, *tb = nxc_tiledn_smem + (a_t ? C::TILE_BYTES : 0);  // one tile per TRANSPOSED operand
  // byte offset of the 16-byte chunk `chunk` of tile row `row` in a swizzled tile
  auto chunk_at = [](int row, int chunk) { return row * (T * ESZ) + ((chunk ^ ((row / V) % CPR)) << 4); };
  // transposed operands: 16-byte vectors along J at fixed i, straight into shared memory
#pragma unroll
  for (int r = 0; r < NPL; r++) {
    const uint32_t il = ry + RP * r, gi = i0 + il, gj = j0 + vx * V;
    if (gi < g.SI && gj < g.SJ) {
      if (a_t) *reinterpret_cast<uint4 *>(ta + chunk_at(il, vx)) =
          __ldg(reinterpret_cast<const uint4 *>(a + base[KA] + (int64_t)gj + (int64_t)gi * g.si[KA]));
      if (b_t) *reinterpret_cast<uint4 *>(tb + chunk_at(il, vx)) =
          __ldg(reinterpret_cast<const uint4 *>(b + base[KB] + (int64_t)gj + (int64_t)gi * g.si[KB]));
    }
  }
  __syncthreads();
  // output: thread (vx, ry) produces, per pass, the W rows jl0 .. jl0 + W - 1 of output vector vx
#pragma unroll 1
  for (int r = 0; r < NPO; r++) {
    const uint32_t jl0 = (ry + RP * r) * W, gi = i0 + vx * V;
    if (gi >= g.SI || j0 + jl0 >= g.SJ) continue;   // SJ is a multiple of V >= W: the W rows are all in or all out
    S1 ea[W][V]; S2 eb[W][V];
    // straight operands: a 16-byte vector along I per output row (stride 1) or one broadcast element
#pragma unroll
    for (int w = 0; w < W; w++) {
      const int64_t gj = (int64_t)j0 + jl0 + w;
      if (!a_t) {
        const S1 *p = a + base[KA] + gj * g.sj[KA];
        if (g.si[KA] == 0) { S1 sv = p[0]; _Pragma("unroll") for (int q = 0; q < V; q++) ea[w][q] = sv; }
        else nxc_load_vec<S1, V>(p + gi, ea[w]);
      }
      if (K::NIN >= 2 && !b_t) {
        const S2 *p = b + base[KB] + gj * g.sj[KB];
        if (g.si[KB] == 0) { S2 sv = p[0]; _Pragma("unroll") for (int q = 0; q < V; q++) eb[w][q] = sv; }
        else nxc_load_vec<S2, V>(p + gi, eb[w]);
      }
    }
    // transposed operands: V word gathers, each word = W adjacent output rows of source row vx*V + q
    if (a_t || b_t) {
#pragma unroll
      for (int q = 0; q < V; q++) {
        const int off = chunk_at(vx * V + q, (int)jl0 / V) + ((int)jl0 % V) * ESZ;
        if (a_t) {
          union { uint32_t u; S1 e[W]; } wa;
          wa.u = *reinterpret_cast<const uint32_t *>(ta + off);
#pragma unroll
          for (int w = 0; w < W; w++) ea[w][q] = wa.e[w];
        }
        if (b_t) {
          union { uint32_t u; S2 e[W]; } wb;
          wb.u = *reinterpret_cast<const uint32_t *>(tb + off);
#pragma unroll
          for (int w = 0; w < W; w++) eb[w][q] = wb.e[w];
        }
      }
    }
#pragma unroll
    for (int w = 0; w < W; w++) {
      S0 vo[V];
      S3 ec[V];
#pragma unroll
      for (int q = 0; q < V; q++) { ec[q] = S3(); if (K::NIN < 2) eb[w][q] = S2(); }
      nxc_run_vec<K, V>(ea[w], eb[w], ec, vo, prm);
      nxc_store_vec<S0, V>(out + base[0] + ((int64_t)j0 + jl0 + w) * g.sj[0] + gi, vo);
    }
  }
}

// Host side: does this plan want the tiled kernel? Picks J = the dim on which some input has
// unit stride while the output's unit stride is on the last dim.
template <class K, bool ENABLED> struct NxcTiledLaunch {
  static bool go(nxc_ctx *, const NxcMapPlan &, typename K::P, nxc_status *) { return false; }
};
template <class K> struct NxcTiledLaunch<K, true> {
  // every access of the vector kernel is a V-element vector: extents, the strides that move
  // between vectors, and the bases must all be V-aligned, and straight operands must run
  // along I with stride 1 (or broadcast along it)
  template <int V> static bool nxc_tiled_vec_ok(const NxcMapPlan &p, int J, uint32_t tmask) {
    constexpr int NOP = K::NIN + 1;
    const int I = p.ndim - 1;
    if (K::NIN > 2 || p.shape[I] % V || p.shape[J] % V) return false;
    const size_t esz[4] = {sizeof(typename K::S0), sizeof(typename K::S1), sizeof(typename K::S2), sizeof(typename K::S3)};
    for (int k = 0; k < NOP; k++) {
      if (esz[k] != esz[0]) return false;
      const bool tr = (tmask >> k) & 1u;
      const int unit = tr ? J : I;
      if (p.stride[k][unit] != 1 && !(p.stride[k][unit] == 0 && !tr)) return false;
      if (!nxc_aligned(p.base[k], esz[k] * V)) return false;
      for (int d = 0; d < p.ndim; d++)
        if (d != unit && p.stride[k][d] % V) return false;
    }
    return true;
  }
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
    if constexpr (sizeof(S0) <= 2 && K::NIN >= 1) {
      if (nxc_tiled_vec_ok<NxcTiledN<K>::V>(p, J, tmask)) {
        typedef NxcTiledN<K> CN;
        constexpr int T = CN::T;
        g.tiles_j = (g.SJ + T - 1) / T;
        g.tiles_i = (g.SI + T - 1) / T;
        g.tiles_i_div = nxc_fastdiv_make(g.tiles_i);
        g.tiles_ij_div = nxc_fastdiv_make(g.tiles_i * g.tiles_j);
        const int64_t nblocks = nrest_total * g.tiles_j * g.tiles_i;
        constexpr int smem_max = K::NIN * CN::TILE_BYTES;
        const int smem = (((tmask >> 1) & 1) + ((tmask >> 2) & 1)) * CN::TILE_BYTES;
        static bool attr_done = false;   // per instantiation; one device per process
        if (!attr_done && smem_max > 48 * 1024) {
          cudaError_t ae = cudaFuncSetAttribute(nxc_map_tiledn_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
          if (ae != cudaSuccess) { *st = nxc_cuda_fail(ctx, ae, "cudaFuncSetAttribute"); return true; }
          attr_done = true;
        }
        nxc_map_tiledn_kernel<K><<<(unsigned)nblocks, 256, smem, ctx->stream>>>(
            (S0 *)p.base[0], (const S1 *)(NOP > 1 ? p.base[1] : p.base[0]), (const S2 *)(NOP > 2 ? p.base[2] : p.base[0]),
            (const S3 *)(NOP > 3 ? p.base[3] : p.base[0]), g, prm);
        ctx->launches++;
        cudaError_t e = cudaPeekAtLastError();
        *st = (e == cudaSuccess) ? NXC_OK : nxc_cuda_fail(ctx, e, "kernel launch");
        return true;
      }
    }
    if (nxc_tiled_vec_ok<NxcTiledV<K>::V>(p, J, tmask)) {
      constexpr int T = NxcTiledV<K>::T;
      g.tiles_j = (g.SJ + T - 1) / T;
      g.tiles_i = (g.SI + T - 1) / T;
      g.tiles_i_div = nxc_fastdiv_make(g.tiles_i);
      g.tiles_ij_div = nxc_fastdiv_make(g.tiles_i * g.tiles_j);
      const int64_t vblocks = nrest_total * g.tiles_j * g.tiles_i;
      nxc_map_tiledv_kernel<K><<<(unsigned)vblocks, 256, 0, ctx->stream>>>(
          (S0 *)p.base[0], (const S1 *)(NOP > 1 ? p.base[1] : p.base[0]), (const S2 *)(NOP > 2 ? p.base[2] : p.base[0]),
          (const S3 *)(NOP > 3 ? p.base[3] : p.base[0]), g, prm);
    } else {
      nxc_map_tiled_kernel<K><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
          (S0 *)p.base[0], (const S1 *)(NOP > 1 ? p.base[1] : p.base[0]), (const S2 *)(NOP > 2 ? p.base[2] : p.base[0]),
          (const S3 *)(NOP > 3 ? p.base[3] : p.base[0]), g, prm);
    }
    ctx->launches++;
    cudaError_t e = cudaPeekAtLastError();
    *st = (e == cudaSuccess) ? NXC_OK : nxc_cuda_fail(ctx, e, "kernel launch");
    return true;
  }
};
template <class K, class = void> struct NxcIsTiled { static constexpr bool v = false; };
template <class K> struct NxcIsTiled<K, typename std::enable_if<K::TILED>::type> { static constexpr bool v = true; };

// ---- launcher ------------------------------------------------------------------------

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
  const int64_t inner = p.shape[od];
  // vector width over the inner dim: the output runs forwards with unit stride, inputs run
  // forwards, backwards (a flipped view) or are broadcast along it
  int vw = 1;
  uint32_t neg = 0;
  if (IPT > 1 && inner % IPT == 0) {
    bool ok = true;
    for (int k = 0; k < NOP && ok; k++) {
      const int64_t s = p.stride[k][od];
      if (k == 0 ? s != 1 : (s != 0 && s != 1 && s != -1)) ok = false;
      if (s == 1 || s == -1) {
        const char *first = p.base[k] - (s == -1 ? (int64_t)(IPT - 1) * (int64_t)esz[k] : 0);
        if (!nxc_aligned(first, esz[k] * IPT)) ok = false;
        for (int d = 0; d < od && ok; d++)
          if (p.stride[k][d] % IPT != 0) ok = false;
        if (s == -1) neg |= 1u << k;
      }
    }
    if (ok) vw = IPT;
  }
  g.ndim = od;
  g.bcast_mask = 0;
  g.neg_mask = vw > 1 ? neg : 0;
  g.ni = inner / vw;
  g.rows = p.total / inner;
  for (int k = 0; k < NOP; k++) {
    g.inner_stride[k] = p.stride[k][od] * vw;
    if (p.stride[k][od] == 0) g.bcast_mask |= 1u << k;
    for (int d = 0; d < od; d++) g.stride[k][d] = p.stride[k][d];
  }
  const bool small = g.rows < 0x7FFFFFFFLL;
  for (int d = 0; d < od; d++) {
    g.shape64[d] = p.shape[d];
    g.div[d] = nxc_fastdiv_make(small ? (uint32_t)p.shape[d] : 1u);
  }
  int txl = 0;
  while (txl < 8 && ((int64_t)1 << txl) < g.ni) txl++;
  g.tx_log2 = txl;
  const int64_t TX = (int64_t)1 << txl, TY = NXC_MAP_THREADS >> txl;
  // long inner dim: 4 items per thread along it; short: 4 rows per thread
  const bool rows_mode = g.ni < 2 * TX;
  g.ui_log2 = rows_mode ? 0 : 2;
  const int64_t per_cta_in = rows_mode ? TX : TX * 4, per_cta_rows = rows_mode ? TY * 4 : TY;
  const int64_t chunks = (g.ni + per_cta_in - 1) / per_cta_in;
  const int64_t rowblocks = (g.rows + per_cta_rows - 1) / per_cta_rows;
  if (chunks >= 0x7FFFFFFFLL || rowblocks >= 0x7FFFFFFFLL || chunks * rowblocks >= 0x7FFFFFFFLL)
    return NXC_ERR_TOO_LARGE;
  g.chunks = (uint32_t)chunks;
  g.chunks_div = nxc_fastdiv_make(g.chunks);
  const unsigned blocks = (unsigned)(chunks * rowblocks);
  if (vw == 1) {
    if (rows_mode) nxc_map_strided_kernel<K, 1, true><<<blocks, NXC_MAP_THREADS, 0, ctx->stream>>>(o, a, b, c, g, prm);
    else nxc_map_strided_kernel<K, 1, false><<<blocks, NXC_MAP_THREADS, 0, ctx->stream>>>(o, a, b, c, g, prm);
  } else {
    if (rows_mode) nxc_map_strided_kernel<K, IPT, true><<<blocks, NXC_MAP_THREADS, 0, ctx->stream>>>(o, a, b, c, g, prm);
    else nxc_map_strided_kernel<K, IPT, false><<<blocks, NXC_MAP_THREADS, 0, ctx->stream>>>(o, a, b, c, g, prm);
  }
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
