// nxc_fold.cuh -- the fold (axis reduction) machinery shared by reduce and
// argreduce.
//
// What is computed is the reference's (nx_c_fold.c:63-101, 110-233;
// nx_c_engine.c:1052-1254): accumulate in the dtype's compute type, store once;
// NaN sticks for float max/min; argmax/argmin keep the FIRST index on ties and
// the first NaN wins. How it is computed is not: the reference parallelises over
// OUTPUT elements only (a full reduction runs on one thread,
// nx_c_engine.c:1166-1168); here the reduced extent itself is split over the
// grid and combined by a second, tiny pass, so a 2^28-element full sum streams
// at HBM rate.
//
// Two kernels cover every layout:
//   row   the fastest-varying reduced dim is the thread axis (TPR threads per
//         output, RPB outputs per block, S splits of the reduced range across
//         blocks). 128-bit loads when that dim has unit stride.
//   lane  a KEPT dim with unit input stride is the thread axis (the reference's
//         "streaming" case, nx_c_engine.c:1108-1149): each thread owns VEC
//         adjacent outputs and walks the reduced rows, TY rows in flight per
//         block, S splits across blockIdx.y.
// Partials of split reductions go to a [S][O] scratch in the accumulator type and
// are folded by nxc_fold_finish_kernel.
#pragma once

#include "nxc_map.cuh"

#define NXC_FOLD_THREADS 256

struct NxcDimList {
  int n;
  NxcFastDiv div[NXC_MAX_NDIM];
  int64_t shape[NXC_MAX_NDIM];
  int64_t in_stride[NXC_MAX_NDIM];
  int64_t out_stride[NXC_MAX_NDIM];
};

// offset of linear index `idx` over the dims (row-major, last dim fastest)
__device__ __forceinline__ void nxc_dims_offset(const NxcDimList &d, int64_t idx, bool small,
                                                int64_t &in_off, int64_t &out_off) {
  in_off = 0;
  out_off = 0;
  if (small) {
    uint32_t r = (uint32_t)idx;
    for (int i = d.n - 1; i >= 0; i--) {
      uint32_t q = nxc_fastdiv(r, d.div[i]);
      uint32_t c = r - q * d.div[i].d;
      in_off += (int64_t)c * d.in_stride[i];
      out_off += (int64_t)c * d.out_stride[i];
      r = q;
    }
  } else {
    int64_t r = idx;
    for (int i = d.n - 1; i >= 0; i--) {
      int64_t q = r / d.shape[i];
      int64_t c = r - q * d.shape[i];
      in_off += c * d.in_stride[i];
      out_off += c * d.out_stride[i];
      r = q;
    }
  }
}

struct NxcRowArgs {
  NxcDimList kept;    // output-indexing dims
  NxcDimList router;  // reduced dims other than the inner one
  int64_t L;          // inner reduced extent (elements)
  int64_t s_inner;    // inner reduced stride (elements)
  int64_t R;          // total reduced extent
  int64_t O;          // outputs
  int64_t chunk;      // reduced elements per split (multiple of the vector width)
  int S;              // splits
  int tpr_log2;       // threads per output row (log2), TPR * RPB == NXC_FOLD_THREADS
  int G;              // rows folded concurrently per thread group (1 or 4)
  int small;          // all linear indices < 2^31
};

struct NxcLaneArgs {
  NxcDimList kept;   // kept dims other than the lane
  NxcDimList red;    // all reduced dims
  int64_t C;         // lane extent (elements); input lane stride is 1
  int64_t lane_out;  // output stride of the lane dim
  int64_t R, O;      // reduced extent, outputs (= prod(kept) * C)
  int64_t chunk;     // reduced rows per split
  int64_t lane_tiles;
  int S;
  int tx_log2;       // threads along lanes (log2); TY = NXC_FOLD_THREADS >> tx_log2
  int small;
};

// ---- warp/block combine helpers -------------------------------------------------
template <class A>
__device__ __forceinline__ A nxc_shfl_xor(const A &v, int mask) {
  constexpr int W = (sizeof(A) + 3) / 4;
  union { A a; uint32_t w[W]; } u, r;
#pragma unroll
  for (int i = 0; i < W; i++) u.w[i] = 0;
  u.a = v;
#pragma unroll
  for (int i = 0; i < W; i++) r.w[i] = __shfl_xor_sync(0xffffffffu, u.w[i], mask);
  return r.a;
}

// N elements whose reduced indices are r0, r0 + rs, ...: policies may fold them as a group
// (argreduce takes the group's extreme first and touches its (value, index) pair only when the
// group beats it); the default is the sequential per-element step.
template <class P, class = void> struct NxcHasMany { static constexpr bool v = false; };
template <class P> struct NxcHasMany<P, typename std::enable_if<P::MANY>::type> { static constexpr bool v = true; };
template <class P, class = void> struct NxcHasWarp { static constexpr bool v = false; };
template <class P> struct NxcHasWarp<P, typename std::enable_if<P::WARP>::type> { static constexpr bool v = true; };
template <class P, int N>
__device__ __forceinline__ void nxc_step_many(typename P::A &acc, const typename P::S (&vals)[N], int64_t r0, int64_t rs) {
  if constexpr (NxcHasMany<P>::v) {
    P::template step_many<N>(acc, vals, r0, rs);
  } else {
#pragma unroll
    for (int i = 0; i < N; i++) P::step(acc, vals[i], r0 + i * rs);
  }
}

// rows in flight per thread group on the short-row path: small accumulators afford 8
template <class P> struct NxcFoldG { static constexpr int v = sizeof(typename P::A) <= 8 ? 8 : 4; };
// short-row path: software-pipeline the row loop (next iteration's loads before this one's fold)?
// Needed for the 8-byte types (see the kernel); it doubles the registers holding loaded values, so
// policies whose fold step is itself heavy (argreduce on 4-byte types) opt out and keep 4 CTAs/SM.
template <class P> struct NxcFoldPipe { static constexpr bool v = true; };
// short-row path: prefer few lanes per row (8 work items each) when rows are plentiful? Pays when the
// per-row epilogue is heavy (argreduce's index vote, the f64 NaN-propagating max); costs the cheap
// f32 sum / max 7 % (a warp's load then spans 4 rows instead of one contiguous 512 bytes).
template <class P> struct NxcFoldFewLanes { static constexpr bool v = false; };

// offsets of output `o`: one kept dim (the usual [rows, R] case) is a multiplication, not a
// decode -- cheap enough to redo at the store instead of keeping per-row offsets in registers
__device__ __forceinline__ void nxc_kept_offset(const NxcDimList &d, int64_t o, bool small, int64_t &in_off,
                                                int64_t &out_off) {
  if (d.n == 0) { in_off = 0; out_off = 0; }
  else if (d.n == 1) { in_off = o * d.in_stride[0]; out_off = o * d.out_stride[0]; }
  else nxc_dims_offset(d, o, small, in_off, out_off);
}

// ---- row kernel ------------------------------------------------------------------
// P: reduction policy with
//   typedef S (input storage), A (accumulator), SO (output storage)
//   static A identity(); static void step(A&, S, int64_t r)  (r increases per accumulator);
//   static A combine(A, A)  (order-free merge of partials);
//   static SO finish(A)
template <class P, int VEC>
__global__ void __launch_bounds__(NXC_FOLD_THREADS)
nxc_fold_row_kernel(const typename P::S *__restrict__ in, typename P::SO *__restrict__ out,
                    typename P::A *__restrict__ scratch, const __grid_constant__ NxcRowArgs a) {
  typedef typename P::S S;
  typedef typename P::A A;
  __shared__ A sm[NXC_FOLD_THREADS];
  const int TPR = 1 << a.tpr_log2;
  const int RPB = NXC_FOLD_THREADS >> a.tpr_log2;
  const int tr = threadIdx.x & (TPR - 1);
  const int row_in_block = threadIdx.x >> a.tpr_log2;
  const int64_t rowblock = (int64_t)blockIdx.x / a.S;
  const int split = (int)((int64_t)blockIdx.x - rowblock * a.S);
  const int64_t o = rowblock * RPB + row_in_block;
  const bool live = o < a.O;
  A acc[4];
#pragma unroll
  for (int i = 0; i < 4; i++) acc[i] = P::identity();
  int64_t in_base = 0, out_off = 0;
  if (live) {
    nxc_dims_offset(a.kept, o, a.small, in_base, out_off);
    const int64_t r0 = (int64_t)split * a.chunk;
    int64_t r1 = r0 + a.chunk;
    if (r1 > a.R) r1 = a.R;
    const int64_t step = (int64_t)TPR * VEC;
    int64_t r = r0 + (int64_t)tr * VEC;
    if (a.router.n == 0) {
      const S *p = in + in_base;
      // 4 independent vector loads in flight per thread
      for (; r + 3 * step < r1; r += 4 * step) {
        S v[4][VEC];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (VEC > 1) nxc_load_vec<S, VEC>(p + (r + u * step), v[u]);
          else v[u][0] = p[(r + u * step) * a.s_inner];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) nxc_step_many<P, VEC>(acc[u], v[u], r + u * step, 1);
      }
      for (; r < r1; r += step) {
        S v[VEC];
        if (VEC > 1) nxc_load_vec<S, VEC>(p + r, v);
        else v[0] = p[r * a.s_inner];
        nxc_step_many<P, VEC>(acc[0], v, r, 1);
      }
    } else {
      for (; r < r1; r += step) {
        // r -> (outer reduced index, inner index)
        int64_t ro = r / a.L;
        int64_t ri = r - ro * a.L;
        int64_t off, dummy;
        nxc_dims_offset(a.router, ro, a.small, off, dummy);
        const S *p = in + in_base + off + ri * a.s_inner;
        S v[VEC];
        if (VEC > 1) nxc_load_vec<S, VEC>(p, v);
        else v[0] = p[0];
        nxc_step_many<P, VEC>(acc[0], v, r, 1);
      }
    }
  }
  A t = P::combine(P::combine(acc[0], acc[1]), P::combine(acc[2], acc[3]));
  // combine across the TPR threads of this row
  if (TPR <= 32) {
    for (int m = TPR >> 1; m > 0; m >>= 1) t = P::combine(t, nxc_shfl_xor(t, m));
  } else {
    for (int m = 16; m > 0; m >>= 1) t = P::combine(t, nxc_shfl_xor(t, m));
    sm[threadIdx.x] = t;
    __syncthreads();
    if (tr == 0) {
      for (int w = 32; w < TPR; w += 32) t = P::combine(t, sm[threadIdx.x + w]);
    }
  }
  if (live && tr == 0) {
    if (a.S == 1) out[out_off] = P::finish(t);
    else scratch[(int64_t)split * a.O + o] = t;
  }
}

// ---- short-row kernel ---------------------------------------------------------------
// Rows shorter than 256 work items, one kept dim (the usual [rows, R] case), no split: one
// thread group of TPR lanes folds G rows at once (rows RPB apart), so G independent vector
// loads are in flight per thread and iteration -- 128 bytes for the 4-byte types, what the
// flat map kernel keeps in flight. Rows are a constant pointer step apart: no per-row decode.
template <class P, int VEC>
__global__ void __launch_bounds__(NXC_FOLD_THREADS)
nxc_fold_short_kernel(const typename P::S *__restrict__ in, typename P::SO *__restrict__ out,
                      const __grid_constant__ NxcRowArgs a) {
  typedef typename P::S S;
  typedef typename P::A A;
  constexpr int G = NxcFoldG<P>::v;
  __shared__ A sm[NXC_FOLD_THREADS];
  const int TPR = 1 << a.tpr_log2;
  const int RPB = NXC_FOLD_THREADS >> a.tpr_log2;
  const int tr = threadIdx.x & (TPR - 1);
  const int64_t o_base = (int64_t)blockIdx.x * ((int64_t)RPB * G) + (threadIdx.x >> a.tpr_log2);
  const int64_t kin = a.kept.n ? a.kept.in_stride[0] : 0, kout = a.kept.n ? a.kept.out_stride[0] : 0;
  // rows of this group that exist: g < n_live
  int n_live = 0;
  if (o_base < a.O) {
    const int64_t left = (a.O - o_base + RPB - 1) / RPB;
    n_live = left < G ? (int)left : G;
  }
  A accg[G];
#pragma unroll
  for (int g = 0; g < G; g++) accg[g] = P::identity();
  const int step = TPR * VEC;
  const int R = (int)a.R;
  const int64_t gstep = (int64_t)RPB * kin;
  const S *p = in + o_base * kin + (VEC > 1 ? (int64_t)tr * VEC : (int64_t)tr * a.s_inner);
  const int64_t pstep = VEC > 1 ? (int64_t)step : (int64_t)step * a.s_inner;
  // Rows past the end re-read this group's last live row and their accumulators are never stored,
  // so loads and steps are unconditional; and the loop is software-pipelined by hand: the G loads
  // of iteration i+1 are issued before iteration i is folded. ptxas otherwise sinks every load of
  // the 8-byte types down to its first use (load, add, load, add ... -- G dependent round trips
  // per iteration, 0.45 of the roofline for f64 rows of 256), and neither unconditional loads nor
  // a warp barrier between the phases talks it out of that.
  const int last_live = n_live - 1;
  int r = tr * VEC;
  if constexpr (!NxcFoldPipe<P>::v) {
    for (; n_live > 0 && r < R; r += step, p += pstep) {
      S v[G][VEC];
#pragma unroll
      for (int g = 0; g < G; g++) {
        const S *pg = p + (int64_t)(g < n_live ? g : last_live) * gstep;
        if (VEC > 1) nxc_load_vec<S, VEC>(pg, v[g]);
        else v[g][0] = *pg;
      }
#pragma unroll
      for (int g = 0; g < G; g++) nxc_step_many<P, VEC>(accg[g], v[g], r, 1);
    }
  } else if (n_live > 0 && r < R) {
    S cur[G][VEC], nxt[G][VEC];
#pragma unroll
    for (int g = 0; g < G; g++) {
      const S *pg = p + (int64_t)(g < n_live ? g : last_live) * gstep;
      if (VEC > 1) nxc_load_vec<S, VEC>(pg, cur[g]);
      else cur[g][0] = *pg;
    }
    for (;;) {
      const int rn = r + step;
      const bool more = rn < R;
      if (more) {
        p += pstep;
#pragma unroll
        for (int g = 0; g < G; g++) {
          const S *pg = p + (int64_t)(g < n_live ? g : last_live) * gstep;
          if (VEC > 1) nxc_load_vec<S, VEC>(pg, nxt[g]);
          else nxt[g][0] = *pg;
        }
      }
#pragma unroll
      for (int g = 0; g < G; g++) nxc_step_many<P, VEC>(accg[g], cur[g], r, 1);
      if (!more) break;
#pragma unroll
      for (int g = 0; g < G; g++)
#pragma unroll
        for (int e = 0; e < VEC; e++) cur[g][e] = nxt[g][e];
      r = rn;
    }
  }
#pragma unroll
  for (int g = 0; g < G; g++) {
    A t = accg[g];
    if (TPR <= 32) {
      if constexpr (NxcHasWarp<P>::v) {
        const unsigned gmask = TPR == 32 ? 0xffffffffu : (((1u << TPR) - 1u) << ((threadIdx.x & 31) & ~(TPR - 1)));
        t = P::warp_combine(t, TPR, gmask);
      } else {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1)
          if (m < TPR) t = P::combine(t, nxc_shfl_xor(t, m));
      }
    } else {
      for (int m = 16; m > 0; m >>= 1) t = P::combine(t, nxc_shfl_xor(t, m));
      __syncthreads();
      sm[threadIdx.x] = t;
      __syncthreads();
      if (tr == 0) for (int w = 32; w < TPR; w += 32) t = P::combine(t, sm[threadIdx.x + w]);
    }
    if (g < n_live && tr == 0) out[(o_base + (int64_t)g * RPB) * kout] = P::finish(t);
  }
}

// ---- lane kernel -----------------------------------------------------------------
template <class P, int VEC>
__global__ void __launch_bounds__(NXC_FOLD_THREADS)
nxc_fold_lane_kernel(const typename P::S *__restrict__ in, typename P::SO *__restrict__ out,
                     typename P::A *__restrict__ scratch, const __grid_constant__ NxcLaneArgs a) {
  typedef typename P::S S;
  typedef typename P::A A;
  __shared__ A sm[NXC_FOLD_THREADS * VEC];
  const int TX = 1 << a.tx_log2;
  const int TY = NXC_FOLD_THREADS >> a.tx_log2;
  const int tx = threadIdx.x & (TX - 1);
  const int ty = threadIdx.x >> a.tx_log2;
  const int64_t ko = (int64_t)blockIdx.x / a.lane_tiles;  // index over the other kept dims
  const int64_t tile = (int64_t)blockIdx.x - ko * a.lane_tiles;
  const int64_t lane0 = (tile * TX + tx) * VEC;
  const bool live = lane0 < a.C;
  const int split = blockIdx.y;
  A acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; j++) acc[j] = P::identity();
  int64_t in_base = 0, out_base = 0;
  nxc_dims_offset(a.kept, ko, a.small, in_base, out_base);
  if (live) {
    const int64_t r0 = (int64_t)split * a.chunk;
    int64_t r1 = r0 + a.chunk;
    if (r1 > a.R) r1 = a.R;
    const S *p = in + in_base + lane0;
    int64_t r = r0 + ty;
    if (a.red.n == 1) {
      const int64_t rs = a.red.in_stride[0];
      constexpr int UL = NxcFoldG<P>::v;  // row loads in flight per thread
      for (; r + (UL - 1) * TY < r1; r += UL * TY) {
        S v[UL][VEC];
#pragma unroll
        for (int u = 0; u < UL; u++) nxc_load_vec<S, VEC>(p + (r + u * TY) * rs, v[u]);
#pragma unroll
        for (int j = 0; j < VEC; j++) {
          S col[UL];
#pragma unroll
          for (int u = 0; u < UL; u++) col[u] = v[u][j];
          nxc_step_many<P, UL>(acc[j], col, r, TY);
        }
      }
      // a few rows left (or a few rows in all: the sum over a batch of 4 after a batched matmul): four in flight
      for (; r + 3 * TY < r1; r += 4 * TY) {
        S v[4][VEC];
#pragma unroll
        for (int u = 0; u < 4; u++) nxc_load_vec<S, VEC>(p + (r + u * TY) * rs, v[u]);
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int j = 0; j < VEC; j++) P::step(acc[j], v[u][j], r + u * TY);
      }
      for (; r < r1; r += TY) {
        S v[VEC];
        nxc_load_vec<S, VEC>(p + r * rs, v);
#pragma unroll
        for (int j = 0; j < VEC; j++) P::step(acc[j], v[j], r);
      }
    } else {
      for (; r < r1; r += TY) {
        int64_t off, dummy;
        nxc_dims_offset(a.red, r, a.small, off, dummy);
        S v[VEC];
        nxc_load_vec<S, VEC>(p + off, v);
#pragma unroll
        for (int j = 0; j < VEC; j++) P::step(acc[j], v[j], r);
      }
    }
  }
  // combine the TY row-walkers of each lane through shared memory
#pragma unroll
  for (int j = 0; j < VEC; j++) sm[(ty * TX + tx) * VEC + j] = acc[j];
  __syncthreads();
  if (ty == 0 && live) {
    for (int y = 1; y < TY; y++)
#pragma unroll
      for (int j = 0; j < VEC; j++) acc[j] = P::combine(acc[j], sm[(y * TX + tx) * VEC + j]);
    const int64_t o_lin = ko * a.C + lane0;
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      if (a.S == 1) out[out_base + (lane0 + j) * a.lane_out] = P::finish(acc[j]);
      else scratch[(int64_t)split * a.O + o_lin + j] = acc[j];
    }
  }
}

// ---- second pass: fold the S partials of each output ------------------------------
struct NxcFinishArgs {
  NxcDimList kept;  // output dims in the order the first pass linearised them
  int64_t O;
  int S;
  int small;
};
template <class P>
__global__ void __launch_bounds__(NXC_FOLD_THREADS)
nxc_fold_finish_kernel(const typename P::A *__restrict__ scratch, typename P::SO *__restrict__ out,
                       const __grid_constant__ NxcFinishArgs a) {
  typedef typename P::A A;
  const int64_t o = (int64_t)blockIdx.x * NXC_FOLD_THREADS + threadIdx.x;
  if (o >= a.O) return;
  A t = scratch[o];
  for (int s = 1; s < a.S; s++) t = P::combine(t, scratch[(int64_t)s * a.O + o]);
  int64_t in_off, out_off;
  nxc_dims_offset(a.kept, o, a.small, in_off, out_off);
  out[out_off] = P::finish(t);
}
// Many partials per output (a kept dim of a few hundred lanes reduced over 2^20 rows leaves
// ~300 of them): a thread walking them one dependent load at a time is latency-bound (60 us for
// the argmax merge); here a WARP owns an output, lanes stride over the partials and merge by
// shuffles -- `combine` is order-free by contract.
template <class P>
__global__ void __launch_bounds__(NXC_FOLD_THREADS)
nxc_fold_finish_warp_kernel(const typename P::A *__restrict__ scratch, typename P::SO *__restrict__ out,
                            const __grid_constant__ NxcFinishArgs a) {
  typedef typename P::A A;
  const int lane = threadIdx.x & 31;
  const int64_t o = (int64_t)blockIdx.x * (NXC_FOLD_THREADS / 32) + (threadIdx.x >> 5);
  if (o >= a.O) return;  // whole warps leave together
  A t = P::identity();
  for (int s = lane; s < a.S; s += 32) t = P::combine(t, scratch[(int64_t)s * a.O + o]);
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) t = P::combine(t, nxc_shfl_xor(t, m));
  if (lane == 0) {
    int64_t in_off, out_off;
    nxc_dims_offset(a.kept, o, a.small, in_off, out_off);
    out[out_off] = P::finish(t);
  }
}
// When there is a single output and many partials (full reductions), one block
// folds them cooperatively.
template <class P>
__global__ void __launch_bounds__(NXC_FOLD_THREADS)
nxc_fold_finish1_kernel(const typename P::A *__restrict__ scratch, typename P::SO *__restrict__ out,
                        int S, int64_t out_off) {
  typedef typename P::A A;
  __shared__ A sm[NXC_FOLD_THREADS / 32];
  A t = P::identity();
  for (int s = threadIdx.x; s < S; s += NXC_FOLD_THREADS) t = P::combine(t, scratch[s]);
  for (int m = 16; m > 0; m >>= 1) t = P::combine(t, nxc_shfl_xor(t, m));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < NXC_FOLD_THREADS / 32; w++) t = P::combine(t, sm[w]);
    out[out_off] = P::finish(t);
  }
}

// ---- host-side plan ---------------------------------------------------------------
struct NxcFoldPlan {
  // coalesced kept dims (in/out element strides) and reduced dims (in strides,
  // sorted by |stride| descending so the last one is the most contiguous)
  int nk, nr;
  int64_t kshape[NXC_MAX_NDIM], k_in[NXC_MAX_NDIM], k_out[NXC_MAX_NDIM];
  int64_t rshape[NXC_MAX_NDIM], r_in[NXC_MAX_NDIM];
  int64_t O, R;
  const char *in_base;
  char *out_base;
};

// Validates like the reference's funnel + driver (nx_c_engine.c:1392-1452,
// 1052-1105) and builds the plan. `out_esize` may differ from `in_esize`
// (argreduce writes int32).
nxc_status nxc_fold_plan(const nxc_tensor *in, const nxc_tensor *out, const int *axes, int n_axes,
                         int64_t in_esize, int64_t out_esize, NxcFoldPlan *p);

static inline void nxc_dimlist_set(NxcDimList &d, int n, const int64_t *shape, const int64_t *in_s,
                                   const int64_t *out_s, bool small) {
  d.n = n;
  for (int i = 0; i < n; i++) {
    d.shape[i] = shape[i];
    d.in_stride[i] = in_s ? in_s[i] : 0;
    d.out_stride[i] = out_s ? out_s[i] : 0;
    d.div[i] = nxc_fastdiv_make(small ? (uint32_t)shape[i] : 1u);
  }
}

// resident CTAs per SM of a kernel (queried once per instantiation): split reductions are sized
// to exactly ONE wave of equal CTAs -- 1280 CTAs on 592 slots measured a third, 16 %-full wave
template <class F>
static inline int nxc_blocks_per_sm(F kernel) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, NXC_FOLD_THREADS, 0) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = 4;
  }
  return n;
}

static inline int nxc_log2_ceil(int64_t x) {
  int l = 0;
  while (((int64_t)1 << l) < x) l++;
  return l;
}

// Launches the right kernel(s) for policy P. `ident_fill` is called by the
// caller beforehand when R == 0.
template <class P>
nxc_status nxc_fold_launch(nxc_ctx *ctx, const NxcFoldPlan &p) {
  typedef typename P::S S;
  typedef typename P::A A;
  typedef typename P::SO SO;
  constexpr int VEC = 16 / (int)sizeof(S) > 0 ? 16 / (int)sizeof(S) : 1;
  const S *in = (const S *)p.in_base;
  SO *out = (SO *)p.out_base;
  const bool small = p.O < 0x7FFFFFFFLL && p.R < 0x7FFFFFFFLL && p.O * (double)p.R < 9.0e18;
  static const int occ_row = nxc_blocks_per_sm(nxc_fold_row_kernel<P, VEC>);
  static const int occ_lane = nxc_blocks_per_sm(nxc_fold_lane_kernel<P, VEC>);

  // lane candidate: kept dim with unit input stride and extent > 1
  int lane = -1;
  for (int j = 0; j < p.nk; j++)
    if (p.kshape[j] > 1 && p.k_in[j] == 1) lane = j;
  int64_t rs_min = INT64_MAX;
  for (int d = 0; d < p.nr; d++) {
    int64_t s = p.r_in[d] < 0 ? -p.r_in[d] : p.r_in[d];
    if (s < rs_min) rs_min = s;
  }
  bool use_lane = lane >= 0 && p.nr >= 1 && rs_min > 1;
  if (use_lane) {
    // vector loads need every row start aligned
    bool ok = (p.kshape[lane] % VEC == 0) && nxc_aligned(in, sizeof(S) * VEC);
    for (int j = 0; j < p.nk && ok; j++)
      if (j != lane && p.k_in[j] % VEC != 0) ok = false;
    for (int d = 0; d < p.nr && ok; d++)
      if (p.r_in[d] % VEC != 0) ok = false;
    const int vec = ok ? VEC : 1;
    NxcLaneArgs a;
    int64_t ks[NXC_MAX_NDIM], ki[NXC_MAX_NDIM], ko[NXC_MAX_NDIM];
    int n = 0;
    for (int j = 0; j < p.nk; j++)
      if (j != lane) { ks[n] = p.kshape[j]; ki[n] = p.k_in[j]; ko[n] = p.k_out[j]; n++; }
    nxc_dimlist_set(a.kept, n, ks, ki, ko, small);
    nxc_dimlist_set(a.red, p.nr, p.rshape, p.r_in, NULL, small);
    a.C = p.kshape[lane];
    a.lane_out = p.k_out[lane];
    a.R = p.R;
    a.O = p.O;
    a.small = small;
    const int64_t lane_items = (a.C + vec - 1) / vec;
    int txl = nxc_log2_ceil(lane_items);
    // 32 threads x 16 B = one 512-byte row segment per warp and 8 row-walkers per lane -- unless there are only a
    // few rows (a [4, 768, 3072] batch sum ran at 1.2 TB/s with half of every CTA idle and a shared-memory
    // merge for 4 values): then fewer walkers, four rows each, and wider CTAs
    int ty_log2 = 3;
    while (ty_log2 > 0 && ((int64_t)4 << ty_log2) > p.R) ty_log2--;
    const int txl_cap = 8 - ty_log2;  // NXC_FOLD_THREADS = 256
    if (txl > txl_cap) txl = txl_cap;
    a.tx_log2 = txl;
    const int TX = 1 << txl, TY = NXC_FOLD_THREADS >> txl;
    a.lane_tiles = (lane_items + TX - 1) / TX;
    const int64_t kother = p.O / a.C;
    const int64_t bx = a.lane_tiles * kother;
    int64_t S_ = 1;
    const int64_t target_blocks = (int64_t)ctx->sm_count * occ_lane;
    // (a few MB in all: a second kernel to merge row splits costs more than the few CTAs of one pass take --
    // the bias-gradient sums of a training step, [256, 768] -> [768], were two 4 us launches each)
    const bool tiny = (double)p.R * (double)p.O * (double)sizeof(S) <= 4.0 * 1048576.0;
    if (bx < target_blocks && !tiny) {
      S_ = target_blocks / bx;  // round down: one wave
      const int64_t max_s = (p.R + (int64_t)TY * NxcFoldG<P>::v - 1) / ((int64_t)TY * NxcFoldG<P>::v);  // >= one unrolled pass per walker
      if (S_ > max_s) S_ = max_s;
      if (S_ > 1024) S_ = 1024;
      if (S_ < 1) S_ = 1;
    }
    a.chunk = (p.R + S_ - 1) / S_;
    S_ = (p.R + a.chunk - 1) / a.chunk;
    a.S = (int)S_;
    A *scr = NULL;
    if (a.S > 1) {
      nxc_status s = nxc_scratch(ctx, sizeof(A) * (size_t)a.S * (size_t)p.O, (void **)&scr);
      if (s) return s;
    }
    dim3 grid((unsigned)bx, (unsigned)a.S);
    if (vec == VEC && VEC > 1)
      nxc_fold_lane_kernel<P, VEC><<<grid, NXC_FOLD_THREADS, 0, ctx->stream>>>(in, out, scr, a);
    else
      nxc_fold_lane_kernel<P, 1><<<grid, NXC_FOLD_THREADS, 0, ctx->stream>>>(in, out, scr, a);
    NXC_LAUNCH_CHECK(ctx);
    if (a.S > 1) {
      NxcFinishArgs f;
      ks[n] = p.kshape[lane]; ki[n] = 0; ko[n] = p.k_out[lane];
      nxc_dimlist_set(f.kept, n + 1, ks, ki, ko, small);
      f.O = p.O; f.S = a.S; f.small = small;
      if (f.S >= 32) {
        const int64_t fb = (p.O + NXC_FOLD_THREADS / 32 - 1) / (NXC_FOLD_THREADS / 32);
        nxc_fold_finish_warp_kernel<P><<<(unsigned)fb, NXC_FOLD_THREADS, 0, ctx->stream>>>(scr, out, f);
      } else {
        const int64_t fb = (p.O + NXC_FOLD_THREADS - 1) / NXC_FOLD_THREADS;
        nxc_fold_finish_kernel<P><<<(unsigned)fb, NXC_FOLD_THREADS, 0, ctx->stream>>>(scr, out, f);
      }
      NXC_LAUNCH_CHECK(ctx);
    }
    return NXC_OK;
  }

  // row path: the last reduced dim (smallest |stride|) is the thread axis
  NxcRowArgs a;
  nxc_dimlist_set(a.kept, p.nk, p.kshape, p.k_in, p.k_out, small);
  if (p.nr == 0) {
    a.L = 1; a.s_inner = 1;
    nxc_dimlist_set(a.router, 0, NULL, NULL, NULL, small);
  } else {
    a.L = p.rshape[p.nr - 1];
    a.s_inner = p.r_in[p.nr - 1];
    nxc_dimlist_set(a.router, p.nr - 1, p.rshape, p.r_in, NULL, small);
  }
  a.R = p.R;
  a.O = p.O;
  a.small = small;
  bool vec_ok = VEC > 1 && a.s_inner == 1 && (a.L % VEC == 0) && nxc_aligned(in, sizeof(S) * VEC);
  for (int j = 0; j < p.nk && vec_ok; j++)
    if (p.k_in[j] % VEC != 0) vec_ok = false;
  for (int d = 0; d + 1 < p.nr && vec_ok; d++)
    if (p.r_in[d] % VEC != 0) vec_ok = false;
  const int vec = vec_ok ? VEC : 1;
  const int64_t items = (p.R + vec - 1) / vec;  // work items per output
  int tl = nxc_log2_ceil(items);
  if (tl > 8) tl = 8;
  // plenty of rows: keep a row inside one warp (shuffle-only combine, no block barrier) and let
  // each lane walk the row; many threads per row only pay off when rows are scarce
  if (tl > 5 && items <= 512 && p.O >= (int64_t)ctx->sm_count * 64) tl = 5;
  // 1- and 2-byte elements: a 16 K-element row is only 16-32 KB, so a block per row spends as long in
  // its barrier-and-shuffle epilogue as in its loads (bf16 / i8 inner-axis sums of [16384, 16384]:
  // 0.75 / 0.49 of the HBM rate). With rows to spare a warp walks the whole row instead.
  // (spreading a row over 2-4 warps to get more blocks for the SMs to balance was tried: 0.72 against
  // 0.77 for bf16 rows of 16384 -- the block barrier in the row epilogue costs more than the tail)
  if (tl > 5 && sizeof(S) <= 2 && items <= 2048 && p.nr <= 1 && p.nk <= 1 && p.O >= (int64_t)ctx->sm_count * 64) tl = 5;
  // ... and with rows to spare, fewer lanes per row, 8 work items each: the per-row epilogue (a
  // shuffle butterfly, for argreduce also the index vote) is paid per thread GROUP, and at 32 lanes
  // per 256-element row it was a quarter of the argmax kernel's instructions. Not below 8 lanes:
  // a group's load stays a full 128-byte line.
  if (NxcFoldFewLanes<P>::v && items <= 512 && p.nr <= 1 && p.nk <= 1 && p.O >= (int64_t)ctx->sm_count * 2048) {
    int want = nxc_log2_ceil(items) - 3;
    if (want < 3) want = 3;
    if (want < tl) tl = want;
  }
  a.tpr_log2 = tl;
  const int RPB = NXC_FOLD_THREADS >> tl;
  a.G = (tl < 8 && p.nr <= 1 && p.nk <= 1 && p.R < 0x7FFFFFFFLL && p.O >= (int64_t)RPB * NxcFoldG<P>::v * ctx->sm_count) ? NxcFoldG<P>::v : 1;
  const int64_t rowblocks = (p.O + (int64_t)RPB * a.G - 1) / ((int64_t)RPB * a.G);
  int64_t S_ = 1;
  const int64_t target_blocks = (int64_t)ctx->sm_count * occ_row;
  if (rowblocks < target_blocks && tl == 8) {
    S_ = target_blocks / rowblocks;  // round down: one wave
    const int64_t per_pass = (int64_t)NXC_FOLD_THREADS * vec * 8;  // >= 8 vectors per thread
    const int64_t max_s = (p.R + per_pass - 1) / per_pass;
    if (S_ > max_s) S_ = max_s;
    if (S_ > 4096) S_ = 4096;
    if (S_ < 1) S_ = 1;
  }
  int64_t chunk = (p.R + S_ - 1) / S_;
  chunk = (chunk + vec - 1) / vec * vec;
  S_ = (p.R + chunk - 1) / chunk;
  if (S_ < 1) S_ = 1;
  a.chunk = chunk;
  a.S = (int)S_;
  A *scr = NULL;
  if (a.S > 1) {
    nxc_status s = nxc_scratch(ctx, sizeof(A) * (size_t)a.S * (size_t)p.O, (void **)&scr);
    if (s) return s;
  }
  const int64_t grid = rowblocks * a.S;
  if (a.G > 1) {
    if (vec > 1) nxc_fold_short_kernel<P, VEC><<<(unsigned)grid, NXC_FOLD_THREADS, 0, ctx->stream>>>(in, out, a);
    else nxc_fold_short_kernel<P, 1><<<(unsigned)grid, NXC_FOLD_THREADS, 0, ctx->stream>>>(in, out, a);
  } else if (vec > 1)
    nxc_fold_row_kernel<P, VEC><<<(unsigned)grid, NXC_FOLD_THREADS, 0, ctx->stream>>>(in, out, scr, a);
  else
    nxc_fold_row_kernel<P, 1><<<(unsigned)grid, NXC_FOLD_THREADS, 0, ctx->stream>>>(in, out, scr, a);
  NXC_LAUNCH_CHECK(ctx);
  if (a.S > 1) {
    if (p.O == 1) {
      int64_t in_off = 0, out_off = 0;
      nxc_fold_finish1_kernel<P><<<1, NXC_FOLD_THREADS, 0, ctx->stream>>>(scr, out, a.S, out_off);
      (void)in_off;
    } else {
      NxcFinishArgs f;
      nxc_dimlist_set(f.kept, p.nk, p.kshape, p.k_in, p.k_out, small);
      f.O = p.O; f.S = a.S; f.small = small;
      if (f.S >= 32) {
        const int64_t fb = (p.O + NXC_FOLD_THREADS / 32 - 1) / (NXC_FOLD_THREADS / 32);
        nxc_fold_finish_warp_kernel<P><<<(unsigned)fb, NXC_FOLD_THREADS, 0, ctx->stream>>>(scr, out, f);
      } else {
        const int64_t fb = (p.O + NXC_FOLD_THREADS - 1) / NXC_FOLD_THREADS;
        nxc_fold_finish_kernel<P><<<(unsigned)fb, NXC_FOLD_THREADS, 0, ctx->stream>>>(scr, out, f);
      }
    }
    NXC_LAUNCH_CHECK(ctx);
  }
  return NXC_OK;
}

template <class P, bool OK> struct NxcMaybeFold {
  static nxc_status go(nxc_ctx *ctx, const NxcFoldPlan &p) { return nxc_fold_launch<P>(ctx, p); }
};
template <class P> struct NxcMaybeFold<P, false> {
  static nxc_status go(nxc_ctx *, const NxcFoldPlan &) { return NXC_ERR_UNSUPPORTED_DTYPE; }
};
