// nxc_fold_mm.cu -- reduce_max / reduce_min instantiations (split from nxc_fold.cu so the
// two halves of the reduction family compile in parallel).
#include "nxc_fold_policy.cuh"

nxc_status nxc_reduce_maxmin(nxc_ctx *ctx, int op, int dt, const NxcFoldPlan &p) {
  nxc_status st = NXC_ERR_UNSUPPORTED_DTYPE;
  switch (op) { NXC_RED_CASE(NXC_RMAX) NXC_RED_CASE(NXC_RMIN) default: break; }
  return st;
}
