// asin acos atan sinh cosh tanh erf (reference: nx_c_map.c:388-456)
#include "nxc_ops.cuh"
#include "nxc_map_groups.cuh"
nxc_status nxc_map1_group_c(nxc_ctx *ctx, int op, int dt, const NxcMapPlan &p) {
  nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
  switch (op) {
    NXC_UN_CASE(NXC_ASIN) NXC_UN_CASE(NXC_ACOS) NXC_UN_CASE(NXC_ATAN)
    NXC_UN_CASE(NXC_SINH) NXC_UN_CASE(NXC_COSH) NXC_UN_CASE(NXC_TANH) NXC_UN_CASE(NXC_ERF)
    default: return NXC_ERR_BAD_OP;
  }
  return st;
}
