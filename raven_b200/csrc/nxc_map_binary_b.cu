// max min pow atan2 xor or and shl shr (reference: nx_c_map.c:594-740)
#include "nxc_ops.cuh"
#include "nxc_map_groups.cuh"
nxc_status nxc_map2_group_b(nxc_ctx *ctx, int op, int dt, const NxcMapPlan &p) {
  nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
  switch (op) {
    NXC_BIN_CASE(NXC_MAX) NXC_BIN_CASE(NXC_MIN) NXC_BIN_CASE(NXC_POW) NXC_BIN_CASE(NXC_ATAN2)
    NXC_BIN_CASE(NXC_XOR) NXC_BIN_CASE(NXC_OR) NXC_BIN_CASE(NXC_AND)
    NXC_BIN_CASE(NXC_SHL) NXC_BIN_CASE(NXC_SHR)
    default: return NXC_ERR_BAD_OP;
  }
  return st;
}
