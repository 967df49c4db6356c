// placeholder: cumulative scans land with the "next" rows of the scope table
#include "nxc_common.cuh"
#define NXC_ERR_NOT_BUILT "operation not implemented in this build"
extern "C" nxc_status nxc_scan(nxc_ctx *, int, const nxc_tensor *, const nxc_tensor *, int) { return NXC_ERR_NOT_BUILT; }
