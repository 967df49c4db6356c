"""oracle/hostview.py -- TEST INFRASTRUCTURE ONLY.

A host-side strided view over a flat numpy storage buffer: the same
{buffer; shape; strides; offset} record the reference passes over its FFI
(reference: packages/nx/lib/backend_c/nx_backend.ml:36-43, nx_c.h:47-61;
strides and offset are in ELEMENTS). Both checkers (oracle/nxo.py, the C
restatement, and oracle/ref.py, the reference's own C compiled unmodified) take
operands in this form, and the GPU parity tests build the device operands from
the very same bytes.

The dtype table mirrors Dtype.Packed.tag order (reference: nx_c.h:130-168,
197-200): tag, numpy storage type, element size (0 for the packed nibble types).
"""
from __future__ import annotations

import numpy as np

# name -> (tag, numpy storage dtype, element bytes)
DTYPES = {
    "f16": (0, np.uint16, 2),
    "f32": (1, np.float32, 4),
    "f64": (2, np.float64, 8),
    "bf16": (3, np.uint16, 2),
    "f8e4m3": (4, np.uint8, 1),
    "f8e5m2": (5, np.uint8, 1),
    "i4": (6, np.uint8, 0),
    "u4": (7, np.uint8, 0),
    "i8": (8, np.int8, 1),
    "u8": (9, np.uint8, 1),
    "i16": (10, np.int16, 2),
    "u16": (11, np.uint16, 2),
    "i32": (12, np.int32, 4),
    "u32": (13, np.uint32, 4),
    "i64": (14, np.int64, 8),
    "u64": (15, np.uint64, 8),
    "c32": (16, np.complex64, 8),
    "c64": (17, np.complex128, 16),
    "bool": (18, np.uint8, 1),
}
TAG_TO_NAME = {v[0]: k for k, v in DTYPES.items()}

FLOATS = ("f16", "f32", "f64", "bf16", "f8e4m3", "f8e5m2")
SINTS = ("i8", "i16", "i32", "i64")
UINTS = ("u8", "u16", "u32", "u64")
INTS = SINTS + UINTS
COMPLEX = ("c32", "c64")
COMPUTE_DTYPES = FLOATS + INTS + COMPLEX + ("bool",)


def tag(dt: str) -> int:
    return DTYPES[dt][0]


def np_storage(dt: str):
    return DTYPES[dt][1]


def esize(dt: str) -> int:
    return DTYPES[dt][2]


def c_strides(shape):
    # Same convention as the reference's Shape.c_contiguous_strides
    # (core/shape.ml:22-33): a zero-extent dim gets stride 0 and counts as
    # extent 1 for the dims to its left.
    n = len(shape)
    st = [0] * n
    if n == 0:
        return st
    st[n - 1] = 0 if int(shape[n - 1]) == 0 else 1
    for i in range(n - 2, -1, -1):
        st[i] = 0 if int(shape[i]) == 0 else st[i + 1] * max(1, int(shape[i + 1]))
    return st


def numel(shape) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


class HostView:
    """storage: 1-D numpy array of the dtype's storage type (shared, mutable)."""

    __slots__ = ("storage", "dtype", "shape", "strides", "offset")

    def __init__(self, storage, dtype, shape, strides=None, offset=0):
        want = np_storage(dtype)
        storage = np.asarray(storage)
        if storage.dtype == np.bool_ and dtype == "bool":
            storage = storage.view(np.uint8)
        if storage.dtype == np.float16 and dtype == "f16":
            storage = storage.view(np.uint16)
        assert storage.ndim == 1 and storage.dtype == want, (storage.dtype, want, dtype)
        assert storage.flags.c_contiguous
        self.storage = storage
        self.dtype = dtype
        self.shape = tuple(int(s) for s in shape)
        self.strides = tuple(int(s) for s in (c_strides(shape) if strides is None else strides))
        assert len(self.shape) == len(self.strides)
        self.offset = int(offset)

    # -- constructors -----------------------------------------------------
    @staticmethod
    def empty(dtype, shape):
        n = numel(shape)
        if esize(dtype) == 0:
            n = (n + 1) // 2
        return HostView(np.zeros(max(n, 0), dtype=np_storage(dtype)), dtype, shape)

    @staticmethod
    def from_array(arr, dtype):
        """Contiguous view over a copy of `arr` (any shape), storage-typed."""
        a = np.ascontiguousarray(arr)
        shape = a.shape
        flat = a.reshape(-1).copy()
        if flat.dtype == np.bool_:
            flat = flat.view(np.uint8)
        if flat.dtype == np.float16:
            flat = flat.view(np.uint16)
        return HostView(flat.astype(np_storage(dtype), copy=False), dtype, shape)

    # -- movement (pure metadata, like core/view.ml) ------------------------
    def permute(self, axes):
        return HostView(self.storage, self.dtype, [self.shape[a] for a in axes],
                        [self.strides[a] for a in axes], self.offset)

    def expand(self, shape):
        if len(self.shape) == 0:  # a scalar expands to any shape (core/view.ml:86-89)
            return HostView(self.storage, self.dtype, shape, [0] * len(shape), self.offset)
        st = [0 if (s == 1 and t != 1) else k for s, t, k in zip(self.shape, shape, self.strides)]
        return HostView(self.storage, self.dtype, shape, st, self.offset)

    def shrink(self, bounds):
        off = self.offset + sum(lo * st for (lo, _), st in zip(bounds, self.strides))
        return HostView(self.storage, self.dtype, [hi - lo for lo, hi in bounds], self.strides, off)

    def flip(self, axes):
        st = list(self.strides)
        off = self.offset
        for a in axes:
            if self.shape[a] > 0:
                off += (self.shape[a] - 1) * st[a]
            st[a] = -st[a]
        return HostView(self.storage, self.dtype, self.shape, st, off)

    def reshape_contig(self, shape):
        assert self.is_contiguous()
        return HostView(self.storage, self.dtype, shape, None, self.offset)

    def is_contiguous(self):
        return tuple(self.strides) == tuple(c_strides(self.shape)) or numel(self.shape) <= 1

    # -- materialisation ----------------------------------------------------
    def numpy(self):
        """Gather the logical elements into a fresh C-contiguous numpy array of
        the storage type (bit patterns for f16/bf16/fp8/bool)."""
        n = numel(self.shape)
        if n == 0:
            return np.zeros(self.shape, dtype=self.storage.dtype)
        idx = np.full(self.shape, self.offset, dtype=np.int64)
        for ax, (s, st) in enumerate(zip(self.shape, self.strides)):
            sh = [1] * len(self.shape)
            sh[ax] = s
            idx = idx + (np.arange(s, dtype=np.int64) * st).reshape(sh)
        return self.storage[idx.reshape(-1)].reshape(self.shape)

    def __repr__(self):
        return f"HostView({self.dtype}, shape={self.shape}, strides={self.strides}, offset={self.offset})"
