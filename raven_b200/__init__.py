"""raven_b200 -- nx-cuda: a B200-native engine behind Nx's `nx.backend` seam.

  raven_b200.backend   host mirror of Nx_backend (Backend_intf.S) over the C ABI
  raven_b200.sharded   leading-axis sharded reduce / argreduce / batch matmul (NCCL)
  raven_b200.csrc      hand-written sm_100a kernels + the C ABI (libnxcuda.so)
"""
from . import dtype  # noqa: F401
from ._lib import Failure, InvalidArgument, LIB_PATH  # noqa: F401
