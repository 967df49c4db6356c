"""ctypes binding of libnxcuda.so (the C ABI in include/nxcuda.h).

This is what an OCaml `external` would bind; the Python mirror goes through the
very same entry points. There is NO CPU fallback: if the shared library is
missing, or no CUDA device is present when a context is created, the product
path raises. The oracle under oracle/ is never imported from here.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnxcuda.so")

NXC_MAX_NDIM = 32


class NxcTensor(ctypes.Structure):
    _fields_ = [
        ("data", ctypes.c_void_p),
        ("dtype", ctypes.c_int32),
        ("ndim", ctypes.c_int32),
        ("shape", ctypes.c_int64 * NXC_MAX_NDIM),
        ("strides", ctypes.c_int64 * NXC_MAX_NDIM),
        ("offset", ctypes.c_int64),
    ]


class InvalidArgument(ValueError):
    """OCaml's Invalid_argument: a precondition violation by the caller."""


class Failure(RuntimeError):
    """OCaml's Failure: unsupported dtype, packed dtype, allocation, CUDA error."""


# every symbol include/nxcuda.h declares: name -> (restype, argtypes)
_P = ctypes.c_void_p
_S = ctypes.c_char_p
_T = ctypes.POINTER(NxcTensor)
SYMBOLS = {
    "nxc_ctx_create": (_S, [ctypes.POINTER(_P)]),
    "nxc_ctx_create_on": (_S, [ctypes.c_int, _P, ctypes.POINTER(_P)]),
    "nxc_ctx_destroy": (None, [_P]),
    "nxc_sync": (_S, [_P]),
    "nxc_stream": (_P, [_P]),
    "nxc_device": (ctypes.c_int, [_P]),
    "nxc_last_error": (_S, [_P]),
    "nxc_launch_count": (ctypes.c_uint64, [_P]),
    "nxc_status_is_invalid_argument": (ctypes.c_int, [_S]),
    "nxc_elem_size": (ctypes.c_int64, [ctypes.c_int]),
    "nxc_set_matmul_mode": (ctypes.c_int, [_P, _S]),
    "nxc_alloc": (_S, [_P, ctypes.c_size_t, ctypes.POINTER(_P)]),
    "nxc_free": (_S, [_P, _P]),
    "nxc_host_alloc": (_S, [_P, ctypes.c_size_t, ctypes.POINTER(_P)]),
    "nxc_host_free": (_S, [_P, _P]),
    "nxc_h2d": (_S, [_P, _P, _P, ctypes.c_size_t]),
    "nxc_d2h": (_S, [_P, _P, _P, ctypes.c_size_t]),
    "nxc_d2h_async": (_S, [_P, _P, _P, ctypes.c_size_t]),
    "nxc_memset": (_S, [_P, _P, ctypes.c_int, ctypes.c_size_t]),
    "nxc_map1": (_S, [_P, ctypes.c_int, _T, _T]),
    "nxc_map2": (_S, [_P, ctypes.c_int, _T, _T, _T]),
    "nxc_cmp": (_S, [_P, ctypes.c_int, _T, _T, _T]),
    "nxc_where": (_S, [_P, _T, _T, _T, _T]),
    "nxc_cast": (_S, [_P, _T, _T]),
    "nxc_copy": (_S, [_P, _T, _T]),
    "nxc_fill": (_S, [_P, _T, _P]),
    "nxc_reduce": (_S, [_P, ctypes.c_int, _T, _T, ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "nxc_argreduce": (_S, [_P, ctypes.c_int, _T, _T, ctypes.c_int]),
    "nxc_scan": (_S, [_P, ctypes.c_int, _T, _T, ctypes.c_int]),
    "nxc_matmul": (_S, [_P, _T, _T, _T]),
    "nxc_cholesky": (_S, [_P, _T, _T, ctypes.c_int]),
    "nxc_triangular_solve": (_S, [_P, _T, _T, _T, ctypes.c_int]),
    "nxc_qr": (_S, [_P, _T, _T, _T, ctypes.c_int]),
    "nxc_eigh": (_S, [_P, _T, _T, _T, ctypes.c_int]),
    "nxc_svd": (_S, [_P, _T, _T, _T, _T]),
    "nxc_eig": (_S, [_P, _T, _T, _T, ctypes.c_int]),
    "nxc_fft": (_S, [_P, _T, _T, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int]),
    "nxc_rfft": (_S, [_P, _T, _T, ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "nxc_irfft": (_S, [_P, _T, _T, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int64]),
    "nxc_pad": (_S, [_P, _T, _T, _P, ctypes.POINTER(ctypes.c_int64)]),
    "nxc_cat": (_S, [_P, _T, ctypes.POINTER(_T), ctypes.c_int, ctypes.c_int]),
    "nxc_gather": (_S, [_P, _T, _T, _T, ctypes.c_int]),
    "nxc_gather_trusted": (_S, [_P, _T, _T, _T, ctypes.c_int]),
    "nxc_scatter": (_S, [_P, _T, _T, _T, ctypes.c_int, ctypes.c_int]),
    "nxc_threefry": (_S, [_P, _T, _T, _T]),
    "nxc_unfold": (_S, [_P, _T, _T, ctypes.c_int] + [ctypes.POINTER(ctypes.c_int64)] * 4),
    "nxc_fold": (_S, [_P, _T, _T, ctypes.c_int] + [ctypes.POINTER(ctypes.c_int64)] * 5),
    "nxc_sort": (_S, [_P, ctypes.c_int, _T, _T, ctypes.c_int, ctypes.c_int]),
    "nxc_dist_unique_id": (_S, [_P]),
    "nxc_dist_init": (_S, [_P, ctypes.c_int, ctypes.c_int, _P]),
    "nxc_dist_finalize": (_S, [_P]),
    "nxc_dist_p2p_enabled": (ctypes.c_int, [_P]),
    "nxc_allreduce": (_S, [_P, _P, ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "nxc_allreduce_async": (_S, [_P, _P, ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "nxc_comm_wait": (_S, [_P]),
    "nxc_allgather": (_S, [_P, _P, _P, ctypes.c_int64]),
    "nxc_argreduce_exchange": (_S, [_P, ctypes.c_int, _T, _T, _T, ctypes.c_int, ctypes.c_int64]),
    "nxc_argreduce_exchange_max_outputs": (ctypes.c_int64, [_P]),
    "nxc_capture_begin": (_S, [_P]),
    "nxc_capture_end": (_S, [_P, ctypes.POINTER(_P)]),
    "nxc_graph_launch": (_S, [_P, _P]),
    "nxc_graph_kernels": (ctypes.c_uint64, [_P]),
    "nxc_graph_arena_bytes": (ctypes.c_size_t, [_P]),
    "nxc_graph_destroy": (None, [_P, _P]),
}

_lib = None


def load():
    """dlopen libnxcuda.so and type every entry point. Raises Failure if the
    extension has not been built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Failure(f"nx-cuda: {LIB_PATH} is missing; build it with __graft_entry__.build(). "
                      "This backend has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(ctx_ptr, op: str, status):
    """Turn a non-NULL status into the exception the reference's funnel raises:
    "<op>: <status>" as Invalid_argument or Failure (nx_c_engine.c:42-52, 1345-1351)."""
    if not status:
        return
    lib = load()
    msg = status.decode()
    detail = ""
    if msg in ("CUDA error", "NCCL error") or msg.startswith(("no CUDA device", "operation not allowed while")):
        d = lib.nxc_last_error(ctx_ptr)
        detail = f" [{d.decode()}]" if d else ""
    text = f"{op}: {msg}{detail}"
    if lib.nxc_status_is_invalid_argument(status):
        raise InvalidArgument(text)
    raise Failure(text)
