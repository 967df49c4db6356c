// cmpeq cmpne cmplt cmple, where, copy, fill
// (reference: nx_c_map.c:747-839; nx_c_move.c:61-84, 191-201). where/copy/fill are
// bit-exact on the STORAGE type, so they are instantiated per element size, not
// per dtype.
#include "nxc_ops.cuh"
#include "nxc_map_groups.cuh"

#define NXC_CMP_CASE(OPC)                                                                   \
  case OPC: {                                                                               \
    NXC_DISPATCH_DTYPE(dt, {                                                                \
      st = NxcMaybeMap<KCmp<OPC, DT>, KCmp<OPC, DT>::O::ok>::go(ctx, p, NxcNoP{}); \
    })                                                                                      \
  } break;

nxc_status nxc_cmp_group(nxc_ctx *ctx, int op, int dt, const NxcMapPlan &p) {
  nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
  switch (op) {
    NXC_CMP_CASE(NXC_CMPEQ) NXC_CMP_CASE(NXC_CMPNE) NXC_CMP_CASE(NXC_CMPLT) NXC_CMP_CASE(NXC_CMPLE)
    default: return NXC_ERR_BAD_OP;
  }
  return st;
}

template <class T> struct KWhere {
  static constexpr int NIN = 3;
  typedef T S0; typedef uint8_t S1; typedef T S2; typedef T S3;
  typedef NxcNoP P;
  __device__ __forceinline__ static S0 run(S1 c, S2 a, S3 b, const P &) { return c ? a : b; }
};
template <class T> struct KCopy {
  static constexpr bool TILED = sizeof(T) <= 8;  // contiguous(transpose x): shared-memory tile transpose
  static constexpr int NIN = 1;
  typedef T S0; typedef T S1; typedef T S2; typedef T S3;
  typedef NxcNoP P;
  __device__ __forceinline__ static S0 run(S1 a, S2, S3, const P &) { return a; }
};
template <class T> struct KFill {
  static constexpr int NIN = 0;
  typedef T S0; typedef T S1; typedef T S2; typedef T S3;
  struct P { T v; };
  __device__ __forceinline__ static S0 run(S1, S2, S3, const P &p) { return p.v; }
};

#define NXC_BY_SIZE(esize, TMPL, ...)                                          \
  switch (esize) {                                                             \
    case 1: return nxc_map_launch<TMPL<uint8_t>>(__VA_ARGS__);                 \
    case 2: return nxc_map_launch<TMPL<uint16_t>>(__VA_ARGS__);                \
    case 4: return nxc_map_launch<TMPL<uint32_t>>(__VA_ARGS__);                \
    case 8: return nxc_map_launch<TMPL<uint2>>(__VA_ARGS__);                   \
    case 16: return nxc_map_launch<TMPL<uint4>>(__VA_ARGS__);                  \
    default: return NXC_ERR_UNSUPPORTED_DTYPE;                                 \
  }

nxc_status nxc_where_group(nxc_ctx *ctx, int esize, const NxcMapPlan &p) {
  NXC_BY_SIZE(esize, KWhere, ctx, p, NxcNoP{})
}
nxc_status nxc_copy_group(nxc_ctx *ctx, int esize, const NxcMapPlan &p) {
  NXC_BY_SIZE(esize, KCopy, ctx, p, NxcNoP{})
}
nxc_status nxc_fill_group(nxc_ctx *ctx, int esize, const NxcMapPlan &p, const void *scalar) {
  switch (esize) {
    case 1: { KFill<uint8_t>::P q; memcpy(&q.v, scalar, 1); return nxc_map_launch<KFill<uint8_t>>(ctx, p, q); }
    case 2: { KFill<uint16_t>::P q; memcpy(&q.v, scalar, 2); return nxc_map_launch<KFill<uint16_t>>(ctx, p, q); }
    case 4: { KFill<uint32_t>::P q; memcpy(&q.v, scalar, 4); return nxc_map_launch<KFill<uint32_t>>(ctx, p, q); }
    case 8: { KFill<uint2>::P q; memcpy(&q.v, scalar, 8); return nxc_map_launch<KFill<uint2>>(ctx, p, q); }
    case 16: { KFill<uint4>::P q; memcpy(&q.v, scalar, 16); return nxc_map_launch<KFill<uint4>>(ctx, p, q); }
    default: return NXC_ERR_UNSUPPORTED_DTYPE;
  }
}
