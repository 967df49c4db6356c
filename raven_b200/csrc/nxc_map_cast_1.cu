// cast matrix rows for sources: NXC_F8E4M3 NXC_F8E5M2 NXC_I8 NXC_U8 NXC_I16 (reference: nx_c_map.c:845-1044)
#include "nxc_ops.cuh"
#include "nxc_cast.cuh"
nxc_status nxc_cast_group1(nxc_ctx *ctx, int src, int dst, const NxcMapPlan &p) {
  switch (src) {
    case NXC_F8E4M3: NXC_CAST_DST_SWITCH(NXC_F8E4M3)
    case NXC_F8E5M2: NXC_CAST_DST_SWITCH(NXC_F8E5M2)
    case NXC_I8: NXC_CAST_DST_SWITCH(NXC_I8)
    case NXC_U8: NXC_CAST_DST_SWITCH(NXC_U8)
    case NXC_I16: NXC_CAST_DST_SWITCH(NXC_I16)
    default: return NXC_ERR_UNSUPPORTED_DTYPE;
  }
}
