// placeholder: threefry lands with the "next" rows of the scope table
#include "nxc_common.cuh"
#define NXC_ERR_NOT_BUILT "operation not implemented in this build"
extern "C" nxc_status nxc_threefry(nxc_ctx *, const nxc_tensor *, const nxc_tensor *, const nxc_tensor *) { return NXC_ERR_NOT_BUILT; }
