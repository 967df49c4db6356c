// cast matrix rows for sources: NXC_F16 NXC_F32 NXC_F64 NXC_BF16 (reference: nx_c_map.c:845-1044)
#include "nxc_ops.cuh"
#include "nxc_cast.cuh"
nxc_status nxc_cast_group0(nxc_ctx *ctx, int src, int dst, const NxcMapPlan &p) {
  switch (src) {
    case NXC_F16: NXC_CAST_DST_SWITCH(NXC_F16)
    case NXC_F32: NXC_CAST_DST_SWITCH(NXC_F32)
    case NXC_F64: NXC_CAST_DST_SWITCH(NXC_F64)
    case NXC_BF16: NXC_CAST_DST_SWITCH(NXC_BF16)
    default: return NXC_ERR_UNSUPPORTED_DTYPE;
  }
}
