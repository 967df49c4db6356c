// add sub mul idiv fdiv mod (reference: nx_c_map.c:499-589)
#include "nxc_ops.cuh"
#include "nxc_map_groups.cuh"
nxc_status nxc_map2_group_a(nxc_ctx *ctx, int op, int dt, const NxcMapPlan &p) {
  nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
  switch (op) {
    NXC_BIN_CASE(NXC_ADD) NXC_BIN_CASE(NXC_SUB) NXC_BIN_CASE(NXC_MUL)
    NXC_BIN_CASE(NXC_IDIV) NXC_BIN_CASE(NXC_FDIV) NXC_BIN_CASE(NXC_MOD)
    default: return NXC_ERR_BAD_OP;
  }
  return st;
}
