// cast matrix rows for sources: NXC_U64 NXC_C32 NXC_C64 NXC_BOOL (reference: nx_c_map.c:845-1044)
#include "nxc_ops.cuh"
#include "nxc_cast.cuh"
nxc_status nxc_cast_group3(nxc_ctx *ctx, int src, int dst, const NxcMapPlan &p) {
  switch (src) {
    case NXC_U64: NXC_CAST_DST_SWITCH(NXC_U64)
    case NXC_C32: NXC_CAST_DST_SWITCH(NXC_C32)
    case NXC_C64: NXC_CAST_DST_SWITCH(NXC_C64)
    case NXC_BOOL: NXC_CAST_DST_SWITCH(NXC_BOOL)
    default: return NXC_ERR_UNSUPPORTED_DTYPE;
  }
}
