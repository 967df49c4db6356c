// sqrt exp log sin cos tan (reference: nx_c_map.c:388-440)
#include "nxc_ops.cuh"
#include "nxc_map_groups.cuh"
nxc_status nxc_map1_group_b(nxc_ctx *ctx, int op, int dt, const NxcMapPlan &p) {
  nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
  switch (op) {
    NXC_UN_CASE(NXC_SQRT) NXC_UN_CASE(NXC_EXP) NXC_UN_CASE(NXC_LOG)
    NXC_UN_CASE(NXC_SIN) NXC_UN_CASE(NXC_COS) NXC_UN_CASE(NXC_TAN)
    default: return NXC_ERR_BAD_OP;
  }
  return st;
}
