// neg recip abs sign trunc ceil floor round (reference: nx_c_map.c:305-383, 460-488)
#include "nxc_ops.cuh"
#include "nxc_map_groups.cuh"
nxc_status nxc_map1_group_a(nxc_ctx *ctx, int op, int dt, const NxcMapPlan &p) {
  nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
  switch (op) {
    NXC_UN_CASE(NXC_NEG) NXC_UN_CASE(NXC_RECIP) NXC_UN_CASE(NXC_ABS) NXC_UN_CASE(NXC_SIGN)
    NXC_UN_CASE(NXC_TRUNC) NXC_UN_CASE(NXC_CEIL) NXC_UN_CASE(NXC_FLOOR) NXC_UN_CASE(NXC_ROUND)
    default: return NXC_ERR_BAD_OP;
  }
  return st;
}
