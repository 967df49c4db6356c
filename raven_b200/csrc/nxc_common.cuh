// nxc_common.cuh -- dtype traits, bit-exact storage<->compute converters, the
// context, and launch helpers shared by every kernel family of libnxcuda.
//
// The dtype semantics implemented here are the reference's (nx_c.h:130-168,
// 345-363; buffer/nx_buffer_stubs.h:73-302): f16/bf16/fp8 compute in float,
// integers wrap on store, bool is 0/1. Integers are computed in their native
// width on the device: every integer op the reference performs in a widened
// 64-bit type and then wraps on store is a ring operation (or an order
// operation on values that fit), so the native-width result is identical.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nxcuda.h"

// ---- status strings (identical text to the reference) ----------------------
#define NXC_ERR_NDIM "ndim exceeds NX_C_MAX_NDIM"
#define NXC_ERR_RANK_MISMATCH "shape and strides rank disagree"
#define NXC_ERR_BAD_KIND "unsupported bigarray kind"
#define NXC_ERR_UNSUPPORTED_DTYPE "dtype not supported for this operation"
#define NXC_ERR_PACKED "packed dtype not supported for this operation"
#define NXC_ERR_SHAPE "shape mismatch"
#define NXC_ERR_EMPTY_REDUCE "reduction over empty axis has no identity"
#define NXC_ERR_ALLOC "out of memory"
#define NXC_ERR_ARGREDUCE_CAP "argreduce axis length exceeds INT32_MAX"
#define NXC_ERR_AXES "reduce axes must be strictly increasing and in range"
#define NXC_ERR_OUT_RANK "output rank inconsistent with the operation"
#define NXC_ERR_AXIS "axis out of range"
#define NXC_ERR_OUT_ALIASED "output has a broadcast (zero) stride"
#define NXC_ERR_ARITY "operand count exceeds NX_C_MAX_OPERANDS"
#define NXC_ERR_DTYPE_MISMATCH "matmul operands must share one dtype"
// new in this engine
#define NXC_ERR_CUDA "CUDA error"
#define NXC_ERR_NO_DEVICE "no CUDA device available (this backend has no CPU fallback)"
#define NXC_ERR_NCCL "NCCL error"
#define NXC_ERR_BAD_OP "unknown operation code"
#define NXC_ERR_TOO_LARGE "iteration space exceeds the launch grid"
#define NXC_ERR_INDEX_DTYPE "indices must be int32"
#define NXC_ERR_INDEX_OOB "index out of bounds for the gathered/scattered axis"
#define NXC_ERR_EXCHANGE_TIMEOUT "peer-memory exchange timed out waiting for a peer rank"
#define NXC_ERR_CAPTURE "operation not allowed while a step is being captured"

struct nxc_ctx {
  int device;
  cudaStream_t stream;
  bool own_stream;
  int sm_count;
  uint64_t launches;
  int matmul_tf32;
  char err[512];
  // scratch for two-pass reductions (grown on demand, stream-ordered)
  void *scratch;
  size_t scratch_bytes;
  // NCCL (dlopen'ed), see nxc_dist.cu
  void *nccl_comm;
  int rank, world;
  // peer-memory mailboxes for small exchanges (CUDA IPC over NVLink), see nxc_dist.cu
  struct nxc_p2p *p2p;
  // TMA descriptor encoder (driver entry point fetched at runtime)
  void *encode_tiled;
  // Side streams (created on first use): pinned host->device and device->host copies run on
  // their own streams so the two PCIe directions and the compute stream overlap; collectives
  // issued through nxc_allreduce_async run on comm_stream under the following kernels.
  cudaStream_t h2d_stream, d2h_stream, comm_stream;
  cudaEvent_t ev_fork, ev_h2d, ev_comm;  // reusable ordering events
  // device buffers with an nxc_d2h_async still reading them: their free is deferred
  struct nxc_pending { void *ptr; cudaEvent_t done; int freed; } *pending;
  int n_pending, cap_pending;
  // Status page: a few words of mapped page-locked host memory that kernels write and the host
  // reads without a copy. Sticky until reported by nxc_sync / nxc_d2h (or the call itself):
  //   [NXC_ST_EXCHANGE] a peer-memory exchange gave up waiting for a peer (nxc_dist.cu)
  //   [NXC_ST_INDEX]    a gather / scatter met an out-of-range index (nxc_move.cu)
  volatile int *hstatus;
  int *dstatus;        // the same page as the device sees it
  int dist_poisoned;   // an exchange timed out: the epochs of the ranks no longer agree
  // Step capture (nxc_capture_begin .. nxc_capture_end, nxc_runtime.cu): while `capturing` is set
  // every launch on `stream` is recorded into a CUDA graph and device memory comes from the graph's
  // own arena instead of the stream-ordered pool.
  struct nxc_graph *capturing;
  struct nxc_graph **graphs;
  int n_graphs, cap_graphs;
  void *saved_scratch;
  size_t saved_scratch_bytes;
  int mm_attr_set;     // the tensor-core GEMM's shared-memory attributes are set on this device
  int nvtx;            // NX_CUDA_NVTX=1: one NVTX range per ABI call (nxc_nvtx.h)
};
enum { NXC_ST_EXCHANGE = 0, NXC_ST_INDEX = 1, NXC_ST_WORDS = 16 };
nxc_status nxc_side_streams(nxc_ctx *ctx);
// reports (and clears) what the status page holds; called after a stream drain
nxc_status nxc_status_page_check(nxc_ctx *ctx);
static inline bool nxc_is_capturing(const nxc_ctx *ctx) { return ctx->capturing != NULL; }
// a call that has to block on the stream (or cannot be replayed) while a step is being captured
nxc_status nxc_capture_refuse(nxc_ctx *ctx, const char *what);

nxc_status nxc_cuda_fail(nxc_ctx *ctx, cudaError_t e, const char *what);
nxc_status nxc_scratch(nxc_ctx *ctx, size_t bytes, void **out);

// ---- tracing: NX_CUDA_NVTX=1 brackets every ABI call in an NVTX range named
// after the entry point, so an nsys / ncu timeline shows the Nx op that owns each kernel. Off by
// default: the cost is then one predictable branch per call. NVTX v3 is header-only (it looks
// for an injected tool library at first use and is a no-op without one).
#include <nvtx3/nvToolsExt.h>
struct NxcTraceScope {
  bool on;
  NxcTraceScope(const nxc_ctx *c, const char *name) : on(c && c->nvtx) { if (on) nvtxRangePushA(name); }
  ~NxcTraceScope() { if (on) nvtxRangePop(); }
};
#define NXC_TRACE(ctx, name) NxcTraceScope nxc_trace_scope_((ctx), (name))

#define NXC_CUDA_TRY(ctx, expr)                                   \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) return nxc_cuda_fail((ctx), _e, #expr); \
  } while (0)

#define NXC_LAUNCH_CHECK(ctx)                                            \
  do {                                                                   \
    (ctx)->launches++;                                                   \
    cudaError_t _e = cudaPeekAtLastError();                              \
    if (_e != cudaSuccess) return nxc_cuda_fail((ctx), _e, "kernel launch"); \
  } while (0)

// ---- dtype classes ---------------------------------------------------------
enum { NXC_CLS_SINT = 1, NXC_CLS_UINT = 2, NXC_CLS_FLOAT = 4, NXC_CLS_COMPLEX = 8,
       NXC_CLS_BOOL = 16, NXC_CLS_PACKED = 32 };

static inline int nxc_dtype_class(int dt) {
  switch (dt) {
    case NXC_F16: case NXC_F32: case NXC_F64: case NXC_BF16: case NXC_F8E4M3: case NXC_F8E5M2:
      return NXC_CLS_FLOAT;
    case NXC_I4: return NXC_CLS_SINT | NXC_CLS_PACKED;
    case NXC_U4: return NXC_CLS_UINT | NXC_CLS_PACKED;
    case NXC_I8: case NXC_I16: case NXC_I32: case NXC_I64: return NXC_CLS_SINT;
    case NXC_U8: case NXC_U16: case NXC_U32: case NXC_U64: return NXC_CLS_UINT;
    case NXC_C32: case NXC_C64: return NXC_CLS_COMPLEX;
    case NXC_BOOL: return NXC_CLS_BOOL;
    default: return 0;
  }
}
static inline bool nxc_is_packed(int dt) { return (nxc_dtype_class(dt) & NXC_CLS_PACKED) != 0; }
static inline bool nxc_valid_dtype(int dt) { return dt >= 0 && dt < NXC_DTYPE_COUNT; }

// ---- complex value types ---------------------------------------------------
struct __align__(8) cf32 { float re, im; };
struct __align__(16) cf64 { double re, im; };

// ---- storage tags: distinct C++ types for dtypes that share a machine type --
struct f16_s { uint16_t b; };
struct bf16_s { uint16_t b; };
struct f8e4_s { uint8_t b; };
struct f8e5_s { uint8_t b; };
struct bool_s { uint8_t b; };

// ---- converters (bit-exact with buffer/nx_buffer_stubs.h) ---------------------
__host__ __device__ inline uint32_t nxc_f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } x; x.f = f; return x.u;
#endif
}
__host__ __device__ inline float nxc_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } x; x.u = u; return x.f;
#endif
}

// half -> float. NaNs are quieted and keep their payload (nx_buffer_stubs.h:150-182).
__device__ inline float nxc_half_to_float(uint16_t h) {
  if ((h & 0x7FFFu) > 0x7C00u)
    return __uint_as_float(((uint32_t)(h & 0x8000u) << 16) | 0x7F800000u |
                           ((uint32_t)(h & 0x3FFu) << 13) | 0x400000u);
  return __half2float(__ushort_as_half(h));
}
// float -> half, RNE; a NaN keeps its top payload bits (nx_buffer_stubs.h:99-148).
__device__ inline uint16_t nxc_float_to_half(float f) {
  uint32_t b = __float_as_uint(f);
  if ((b & 0x7FFFFFFFu) > 0x7F800000u) {
    uint16_t r = (uint16_t)(0x7C00u + ((b & 0x007FFFFFu) >> 13));
    r += (r == 0x7C00u);
    return (uint16_t)(((b & 0x80000000u) >> 16) + r);
  }
  return __half_as_ushort(__float2half_rn(f));
}
__device__ inline float nxc_bf16_to_float(uint16_t h) { return __uint_as_float((uint32_t)h << 16); }
// float -> bfloat16, RNE, NaN quieted keeping the sign (nx_buffer_stubs.h:73-86).
__device__ inline uint16_t nxc_float_to_bf16(float f) {
  uint32_t b = __float_as_uint(f);
  if ((b & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((b >> 16) | 0x0040u);
  return (uint16_t)((b + (((b >> 16) & 1u) + 0x7FFFu)) >> 16);
}
// Generic small-float encoder: EB exponent bits, MB mantissa bits, RNE with
// subnormals. `maxbits` is the first magnitude pattern that means overflow and
// `ovf` what overflow turns into (e4m3fn: NaN 0x7F; e5m2: Inf 0x7C)
// (nx_buffer_stubs.h:189-223, 244-278).
template <int EB, int MB, int BIAS, uint32_t MAXBITS, uint32_t OVF, uint32_t NANB, uint32_t INFB>
__device__ inline uint8_t nxc_float_to_fp8(float f) {
  uint32_t b = __float_as_uint(f);
  uint32_t sign = (b >> 31) << 7;
  uint32_t mag = b & 0x7FFFFFFFu;
  if (mag > 0x7F800000u) return (uint8_t)NANB;          // NaN: sign dropped
  if (mag == 0x7F800000u) return (uint8_t)(sign | INFB);
  int exp = (int)((b >> 23) & 0xFF) - 127;
  const int emin = 1 - BIAS;
  if (exp >= emin) {
    uint32_t sig = b & 0x7FFFFFu;
    const int sh = 23 - MB;
    uint32_t q = sig >> sh;
    uint32_t rem = sig & ((1u << sh) - 1u);
    uint32_t half = 1u << (sh - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    uint32_t bits = ((uint32_t)(exp + BIAS) << MB) + q;
    if (bits >= MAXBITS) return (uint8_t)(sign | OVF);
    return (uint8_t)(sign | bits);
  }
  uint32_t sig = (b & 0x7FFFFFu) | 0x800000u;
  int shift = (23 - MB) + (emin - exp);
  if (shift > 24) return (uint8_t)sign;
  uint32_t q = sig >> shift;
  uint32_t rem = sig & ((1u << shift) - 1u);
  uint32_t half = 1u << (shift - 1);
  if (rem > half || (rem == half && (q & 1u))) q++;
  return (uint8_t)(sign | q);
}
__device__ inline uint8_t nxc_float_to_e4m3(float f) {
  return nxc_float_to_fp8<4, 3, 7, 0x7Fu, 0x7Fu, 0x7Fu, 0x7Fu>(f);
}
__device__ inline uint8_t nxc_float_to_e5m2(float f) {
  return nxc_float_to_fp8<5, 2, 15, 0x7Cu, 0x7Cu, 0x7Fu, 0x7Cu>(f);
}
// e4m3fn -> float: no infinities, S.1111.111 is (positive, quiet) NaN
// (nx_buffer_stubs.h:225-240).
__device__ inline float nxc_e4m3_to_float(uint8_t v) {
  uint32_t e = (v >> 3) & 0xFu, m = v & 7u;
  if (e == 0xFu && m == 7u) return __uint_as_float(0x7FC00000u);
  float r = (e == 0) ? ldexpf((float)m, -9) : ldexpf(1.0f + (float)m * 0.125f, (int)e - 7);
  return (v & 0x80u) ? -r : r;
}
__device__ inline float nxc_e5m2_to_float(uint8_t v) {
  uint32_t e = (v >> 2) & 0x1Fu, m = v & 3u;
  if (e == 0x1Fu) {
    if (m == 0) return (v & 0x80u) ? -INFINITY : INFINITY;
    return __uint_as_float(0x7FC00000u);
  }
  float r = (e == 0) ? ldexpf((float)m * 0.25f, -14) : ldexpf(1.0f + (float)m * 0.25f, (int)e - 15);
  return (v & 0x80u) ? -r : r;
}

// ---- dtype traits ------------------------------------------------------------
// S = storage type, C = compute type, cls = category.
template <int DT> struct DT_;
#define NXC_DEF_DT(tag, S_, C_, cls_, LD, ST)                         \
  template <> struct DT_<tag> {                                       \
    typedef S_ S;                                                     \
    typedef C_ C;                                                     \
    static constexpr int cls = cls_;                                  \
    static constexpr int tag_ = tag;                                  \
    __device__ static inline C ld(S s) { return LD; }                 \
    __device__ static inline S st(C v) { return ST; }                 \
  };
NXC_DEF_DT(NXC_F16, f16_s, float, NXC_CLS_FLOAT, nxc_half_to_float(s.b), (f16_s{nxc_float_to_half(v)}))
NXC_DEF_DT(NXC_F32, float, float, NXC_CLS_FLOAT, s, v)
NXC_DEF_DT(NXC_F64, double, double, NXC_CLS_FLOAT, s, v)
NXC_DEF_DT(NXC_BF16, bf16_s, float, NXC_CLS_FLOAT, nxc_bf16_to_float(s.b), (bf16_s{nxc_float_to_bf16(v)}))
NXC_DEF_DT(NXC_F8E4M3, f8e4_s, float, NXC_CLS_FLOAT, nxc_e4m3_to_float(s.b), (f8e4_s{nxc_float_to_e4m3(v)}))
NXC_DEF_DT(NXC_F8E5M2, f8e5_s, float, NXC_CLS_FLOAT, nxc_e5m2_to_float(s.b), (f8e5_s{nxc_float_to_e5m2(v)}))
NXC_DEF_DT(NXC_I8, int8_t, int32_t, NXC_CLS_SINT, (int32_t)s, (int8_t)v)
NXC_DEF_DT(NXC_U8, uint8_t, uint32_t, NXC_CLS_UINT, (uint32_t)s, (uint8_t)v)
NXC_DEF_DT(NXC_I16, int16_t, int32_t, NXC_CLS_SINT, (int32_t)s, (int16_t)v)
NXC_DEF_DT(NXC_U16, uint16_t, uint32_t, NXC_CLS_UINT, (uint32_t)s, (uint16_t)v)
NXC_DEF_DT(NXC_I32, int32_t, int32_t, NXC_CLS_SINT, s, v)
NXC_DEF_DT(NXC_U32, uint32_t, uint32_t, NXC_CLS_UINT, s, v)
NXC_DEF_DT(NXC_I64, int64_t, int64_t, NXC_CLS_SINT, s, v)
NXC_DEF_DT(NXC_U64, uint64_t, uint64_t, NXC_CLS_UINT, s, v)
NXC_DEF_DT(NXC_C32, cf32, cf32, NXC_CLS_COMPLEX, s, v)
NXC_DEF_DT(NXC_C64, cf64, cf64, NXC_CLS_COMPLEX, s, v)
NXC_DEF_DT(NXC_BOOL, bool_s, uint32_t, NXC_CLS_BOOL, (uint32_t)(s.b != 0), (bool_s{(uint8_t)(v != 0)}))
#undef NXC_DEF_DT

// ---- packed stores of the 16-bit float types ------------------------------------------------------
// N compute-type (float) results -> N storage values, two per hardware convert
// (cvt.rn.{bf16x2,f16x2}.f32: one instruction per PAIR instead of ~9 integer operations per
// element for the software round-to-nearest-even). The hardware turns every NaN into the
// canonical 0x7FFF where the reference keeps sign and payload (nx_buffer_stubs.h:73-86, 99-148),
// so one `setp.nan` per pair collects "was there a NaN at all" and a thread that saw one redoes
// its whole vector with the exact scalar converters -- a branch that is never taken on real data.
// Measured on B200 over all 2^32 float patterns (scratch/cvt_probe.cu): the hardware converts
// agree with the reference's bit for bit on every non-NaN input.
template <int DT> struct NxcPack16 { static constexpr bool v = (DT == NXC_F16 || DT == NXC_BF16); };
template <int DT> __device__ __forceinline__ uint32_t nxc_cvt_pair16(float lo, float hi) {
  uint32_t r;
  if (DT == NXC_BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t nxc_pair_has_nan(float a, float b) {
  uint32_t r;
  asm("{ .reg .pred q; setp.nan.f32 q, %1, %2; selp.u32 %0, 1, 0, q; }" : "=r"(r) : "f"(a), "f"(b));
  return r;
}
template <int DT, int N>
__device__ __forceinline__ void nxc_pack16(const float (&v)[N], typename DT_<DT>::S (&o)[N]) {
  static_assert(N % 2 == 0, "pairs");
  uint32_t w[N / 2], nan = 0;
#pragma unroll
  for (int j = 0; j < N / 2; j++) {
    w[j] = nxc_cvt_pair16<DT>(v[2 * j], v[2 * j + 1]);
    nan |= nxc_pair_has_nan(v[2 * j], v[2 * j + 1]);
  }
  if (nan) {
#pragma unroll
    for (int i = 0; i < N; i++) o[i] = DT_<DT>::st(v[i]);
  } else {
#pragma unroll
    for (int j = 0; j < N / 2; j++) {
      o[2 * j].b = (uint16_t)(w[j] & 0xFFFFu);
      o[2 * j + 1].b = (uint16_t)(w[j] >> 16);
    }
  }
}

// N storage values -> N compute values. For f16 the hardware widening convert is exact on every
// non-NaN pattern but canonicalises NaNs, where the reference quiets them and keeps the payload
// (nx_buffer_stubs.h:150-182; measured: all 2046 NaN patterns differ, no other): convert
// everything in hardware, test the RESULTS for NaN a pair at a time, and redo the vector with the
// exact scalar converter only in a thread that met one. Every other dtype: N scalar loads.
template <int DT, int N>
__device__ __forceinline__ void nxc_ld_many(const typename DT_<DT>::S (&s)[N], typename DT_<DT>::C (&c)[N]) {
  if constexpr (DT == NXC_F16 && N % 2 == 0) {
    uint32_t nan = 0;
#pragma unroll
    for (int i = 0; i < N; i++) c[i] = __half2float(__ushort_as_half(s[i].b));
#pragma unroll
    for (int j = 0; j < N / 2; j++) nan |= nxc_pair_has_nan(c[2 * j], c[2 * j + 1]);
    if (nan) {
#pragma unroll
      for (int i = 0; i < N; i++) c[i] = nxc_half_to_float(s[i].b);
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; i++) c[i] = DT_<DT>::ld(s[i]);
  }
}

// Dispatch a runtime dtype tag to a template instantiation over the 17 compute
// dtypes. BODY sees `DT` as a constexpr int.
#define NXC_DISPATCH_DTYPE(dt, ...)                                             \
  switch (dt) {                                                                  \
    case NXC_F16: { constexpr int DT = NXC_F16; __VA_ARGS__ } break;                    \
    case NXC_F32: { constexpr int DT = NXC_F32; __VA_ARGS__ } break;                    \
    case NXC_F64: { constexpr int DT = NXC_F64; __VA_ARGS__ } break;                    \
    case NXC_BF16: { constexpr int DT = NXC_BF16; __VA_ARGS__ } break;                  \
    case NXC_F8E4M3: { constexpr int DT = NXC_F8E4M3; __VA_ARGS__ } break;              \
    case NXC_F8E5M2: { constexpr int DT = NXC_F8E5M2; __VA_ARGS__ } break;              \
    case NXC_I8: { constexpr int DT = NXC_I8; __VA_ARGS__ } break;                      \
    case NXC_U8: { constexpr int DT = NXC_U8; __VA_ARGS__ } break;                      \
    case NXC_I16: { constexpr int DT = NXC_I16; __VA_ARGS__ } break;                    \
    case NXC_U16: { constexpr int DT = NXC_U16; __VA_ARGS__ } break;                    \
    case NXC_I32: { constexpr int DT = NXC_I32; __VA_ARGS__ } break;                    \
    case NXC_U32: { constexpr int DT = NXC_U32; __VA_ARGS__ } break;                    \
    case NXC_I64: { constexpr int DT = NXC_I64; __VA_ARGS__ } break;                    \
    case NXC_U64: { constexpr int DT = NXC_U64; __VA_ARGS__ } break;                    \
    case NXC_C32: { constexpr int DT = NXC_C32; __VA_ARGS__ } break;                    \
    case NXC_C64: { constexpr int DT = NXC_C64; __VA_ARGS__ } break;                    \
    case NXC_BOOL: { constexpr int DT = NXC_BOOL; __VA_ARGS__ } break;                  \
    default: break;                                                              \
  }

// ---- host-side operand validation (reference: nx_c.h:420-434) ------------------
static inline nxc_status nxc_check_tensor(const nxc_tensor *t) {
  if (t->ndim < 0 || t->ndim > NXC_MAX_NDIM) return NXC_ERR_NDIM;
  if (!nxc_valid_dtype(t->dtype)) return NXC_ERR_BAD_KIND;
  return NXC_OK;
}
static inline int64_t nxc_numel(const nxc_tensor *t) {
  int64_t n = 1;
  for (int i = 0; i < t->ndim; i++) n *= t->shape[i];
  return n;
}
