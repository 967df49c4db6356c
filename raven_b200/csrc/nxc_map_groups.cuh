// nxc_map_groups.cuh -- how the map family is split over translation units so
// the ~1000 kernel instantiations compile in parallel. Each group function
// dispatches (op, dtype) to one nxc_map_launch instantiation, or reports the
// reference's "dtype not supported" status for a NULL slot in its tables
// (reference: nx_c_map.c:224-294, nx_c_engine.c:837-839).
#pragma once
#include "nxc_map.cuh"

nxc_status nxc_map1_group_a(nxc_ctx *ctx, int op, int dt, const NxcMapPlan &p);  // neg recip abs sign + rounding
nxc_status nxc_map1_group_b(nxc_ctx *ctx, int op, int dt, const NxcMapPlan &p);  // sqrt exp log sin cos tan
nxc_status nxc_map1_group_c(nxc_ctx *ctx, int op, int dt, const NxcMapPlan &p);  // asin..tanh erf
nxc_status nxc_map2_group_a(nxc_ctx *ctx, int op, int dt, const NxcMapPlan &p);  // add sub mul idiv fdiv mod
nxc_status nxc_map2_group_b(nxc_ctx *ctx, int op, int dt, const NxcMapPlan &p);  // max min pow atan2 xor or and shl shr
nxc_status nxc_cmp_group(nxc_ctx *ctx, int op, int dt, const NxcMapPlan &p);
nxc_status nxc_where_group(nxc_ctx *ctx, int esize, const NxcMapPlan &p);
nxc_status nxc_copy_group(nxc_ctx *ctx, int esize, const NxcMapPlan &p);
nxc_status nxc_fill_group(nxc_ctx *ctx, int esize, const NxcMapPlan &p, const void *scalar);
nxc_status nxc_cast_group(nxc_ctx *ctx, int src, int dst, const NxcMapPlan &p);
nxc_status nxc_cast_group0(nxc_ctx *ctx, int src, int dst, const NxcMapPlan &p);  // src f16 f32 f64 bf16
nxc_status nxc_cast_group1(nxc_ctx *ctx, int src, int dst, const NxcMapPlan &p);  // src fp8s i8 u8 i16
nxc_status nxc_cast_group2(nxc_ctx *ctx, int src, int dst, const NxcMapPlan &p);  // src u16 i32 u32 i64
nxc_status nxc_cast_group3(nxc_ctx *ctx, int src, int dst, const NxcMapPlan &p);  // src u64 c32 c64 bool

#define NXC_UN_CASE(OPC)                                                                  \
  case OPC: {                                                                             \
    NXC_DISPATCH_DTYPE(dt, {                                                              \
      st = NxcMaybeMap<KUn<OPC, DT>, KUn<OPC, DT>::O::ok>::go(ctx, p, NxcNoP{}); \
    })                                                                                    \
  } break;
#define NXC_BIN_CASE(OPC)                                                                   \
  case OPC: {                                                                               \
    NXC_DISPATCH_DTYPE(dt, {                                                                \
      st = NxcMaybeMap<KBin<OPC, DT>, KBin<OPC, DT>::O::ok>::go(ctx, p, NxcNoP{}); \
    })                                                                                      \
  } break;
