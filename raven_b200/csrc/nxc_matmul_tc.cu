// placeholder until the tcgen05 path lands (next commit)
#include "nxc_matmul.cuh"
nxc_status nxc_matmul_tc(nxc_ctx *, const NxcMatmulProblem &) { return NXC_MM_TC_DECLINED; }
