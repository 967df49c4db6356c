// cast matrix rows for sources: NXC_U16 NXC_I32 NXC_U32 NXC_I64 (reference: nx_c_map.c:845-1044)
#include "nxc_ops.cuh"
#include "nxc_cast.cuh"
nxc_status nxc_cast_group2(nxc_ctx *ctx, int src, int dst, const NxcMapPlan &p) {
  switch (src) {
    case NXC_U16: NXC_CAST_DST_SWITCH(NXC_U16)
    case NXC_I32: NXC_CAST_DST_SWITCH(NXC_I32)
    case NXC_U32: NXC_CAST_DST_SWITCH(NXC_U32)
    case NXC_I64: NXC_CAST_DST_SWITCH(NXC_I64)
    default: return NXC_ERR_UNSUPPORTED_DTYPE;
  }
}
