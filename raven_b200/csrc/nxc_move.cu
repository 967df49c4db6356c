// placeholder: pad / cat / gather / scatter land with the "next" rows of the scope table
#include "nxc_common.cuh"
#define NXC_ERR_NOT_BUILT "operation not implemented in this build"
extern "C" nxc_status nxc_pad(nxc_ctx *, const nxc_tensor *, const nxc_tensor *, const void *, const int64_t *) { return NXC_ERR_NOT_BUILT; }
extern "C" nxc_status nxc_cat(nxc_ctx *, const nxc_tensor *, const nxc_tensor *const *, int, int) { return NXC_ERR_NOT_BUILT; }
extern "C" nxc_status nxc_gather(nxc_ctx *, const nxc_tensor *, const nxc_tensor *, const nxc_tensor *, int) { return NXC_ERR_NOT_BUILT; }
extern "C" nxc_status nxc_scatter(nxc_ctx *, const nxc_tensor *, const nxc_tensor *, const nxc_tensor *, int, int) { return NXC_ERR_NOT_BUILT; }
