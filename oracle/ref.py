"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY.

Drives the reference's own C backend (packages/nx/lib/backend_c/nx_c_*.c,
compiled UNMODIFIED into oracle/_ref/libnxref.so by oracle/Makefile) from
Python: builds OCaml-layout operand records around numpy storage and calls the
real `caml_nx_c_*` CAMLprim stubs, so the reference's own funnel, coalescing,
streaming-fold selection, thread policy and error strings are what answers.

Record layout (reference: nx_c.h:47-61, 420-434): a block of >= 4 fields
[bigarray; shape int array; strides int array; offset], strides/offset in
ELEMENTS. Bigarray = custom block whose struct caml_ba_array starts at field 1;
extended kinds live in flag bits 16-23 (reference: buffer/nx_buffer_stubs.h:28-68).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module. The product path never does.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .hostview import HostView, numel

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libnxref.so")

# bigarray kind per dtype: (standard kind in flags&0xFF, extended kind in bits 16-23)
_KIND = {
    "f32": (0, 0), "f64": (1, 0), "i8": (2, 0), "u8": (3, 0), "i16": (4, 0),
    "u16": (5, 0), "i32": (6, 0), "i64": (7, 0), "c32": (10, 0), "c64": (11, 0),
    "f16": (13, 0),
    # extended kinds keep a plausible base kind of the same width underneath
    "bf16": (5, 14), "bool": (3, 15), "i4": (3, 16), "u4": (3, 17),
    "f8e4m3": (3, 18), "f8e5m2": (3, 19), "u32": (6, 20), "u64": (7, 21),
}

_lib = None


class RefError(Exception):
    """The stub raised. kind is 'Failure' or 'Invalid_argument'."""

    def __init__(self, kind, msg):
        super().__init__(f"{kind}: {msg}")
        self.kind = kind
        self.msg = msg


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.nxref_invoke.restype = ctypes.c_int
        _lib.nxref_invoke.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _lib.nxref_last_message.restype = ctypes.c_char_p
    return _lib


def val_int(n: int) -> int:
    return ((int(n) << 1) | 1) & 0xFFFFFFFFFFFFFFFF


class _Arena:
    """Keeps every ctypes block alive for the duration of one call."""

    def __init__(self):
        self.keep = []

    def block(self, fields):
        n = len(fields)
        words = (ctypes.c_uint64 * (n + 1))()
        words[0] = n << 10
        for i, f in enumerate(fields):
            words[i + 1] = f & 0xFFFFFFFFFFFFFFFF
        self.keep.append(words)
        return ctypes.addressof(words) + 8

    def int_array(self, xs):
        return self.block([val_int(x) for x in xs])

    def bigarray(self, hv: HostView):
        kind, ext = _KIND[hv.dtype]
        flags = kind | (ext << 16)
        data = hv.storage.ctypes.data if hv.storage.size else 0
        # [ops][data][num_dims][flags][proxy][dim0]
        return self.block([0, data, 1, flags, 0, hv.storage.size])

    def tensor(self, hv: HostView):
        self.keep.append(hv.storage)
        return self.block([self.bigarray(hv), self.int_array(hv.shape),
                           self.int_array(hv.strides), val_int(hv.offset), val_int(0), val_int(0)])


def call(stub: str, *args):
    """args: HostView | int | bool | list[int] | list[HostView]."""
    L = lib()
    fn = ctypes.cast(getattr(L, "caml_nx_c_" + stub), ctypes.c_void_p)
    ar = _Arena()
    vals = []
    for a in args:
        if isinstance(a, HostView):
            vals.append(ar.tensor(a))
        elif isinstance(a, (bool, np.bool_)):
            vals.append(val_int(1 if a else 0))
        elif isinstance(a, (int, np.integer)):
            vals.append(val_int(a))
        elif isinstance(a, (list, tuple)):
            if len(a) and isinstance(a[0], HostView):
                vals.append(ar.block([ar.tensor(x) for x in a]))
            else:
                vals.append(ar.int_array(a))
        else:
            raise TypeError(type(a))
    arr = (ctypes.c_uint64 * len(vals))(*[v & 0xFFFFFFFFFFFFFFFF for v in vals])
    rc = L.nxref_invoke(fn, len(vals), ctypes.cast(arr, ctypes.c_void_p))
    if rc == 1:
        raise RefError("Failure", L.nxref_last_message().decode())
    if rc == 2:
        raise RefError("Invalid_argument", L.nxref_last_message().decode())
    if rc != 0:
        raise RuntimeError(f"nxref_invoke: bad arity for {stub}")


# ---- the veneer, restated (reference: backend_c/nx_backend.ml:170-500) ------

UNARY = ("neg recip abs sign sqrt exp log sin cos tan asin acos atan sinh cosh tanh "
         "trunc ceil floor round erf").split()
BINARY = "add sub mul idiv fdiv mod pow atan2 max min xor or and".split()
CMP = "cmpeq cmpne cmplt cmple".split()


def unary(op, x: HostView) -> HostView:
    out = HostView.empty(x.dtype, x.shape)
    call(op, out, x)
    return out


def binary(op, x: HostView, y: HostView) -> HostView:
    out = HostView.empty(x.dtype, x.shape)
    call(op, out, x, y)
    return out


def compare(op, x, y) -> HostView:
    out = HostView.empty("bool", x.shape)
    call(op, out, x, y)
    return out


def where(c, a, b) -> HostView:
    out = HostView.empty(a.dtype, a.shape)
    call("where", out, c, a, b)
    return out


def cast(x, dtype) -> HostView:
    out = HostView.empty(dtype, x.shape)
    call("cast", out, x)
    return out


def copy(x) -> HostView:
    out = HostView.empty(x.dtype, x.shape)
    call("copy", out, x)
    return out


def assign(dst, src) -> None:
    call("copy", dst, src)


def reduce(op, x: HostView, axes) -> HostView:
    axes = sorted(int(a) for a in axes)
    if op in ("max", "min"):
        for ax in axes:
            if x.shape[ax] == 0:
                raise RefError("Invalid_argument",
                               f"reduce_{op}: reduction over an empty axis has no identity")
    out_shape = [d for i, d in enumerate(x.shape) if i not in axes]
    out = HostView.empty(x.dtype, out_shape)
    call("reduce_" + op, out, x, axes)
    return out


def argreduce(op, x: HostView, axis: int, keepdims: bool = False) -> HostView:
    if x.shape[axis] == 0:
        raise RefError("Invalid_argument", f"{op}: argument reduction over an empty axis")
    if keepdims:
        out_shape = [1 if i == axis else d for i, d in enumerate(x.shape)]
    else:
        out_shape = [d for i, d in enumerate(x.shape) if i != axis]
    out = HostView.empty("i32", out_shape)
    call(op, out, x, int(axis))
    return out


def scan(op, x: HostView, axis: int) -> HostView:
    out = HostView.empty(x.dtype, x.shape)
    call({"sum": "cumsum", "prod": "cumprod", "max": "cummax", "min": "cummin"}[op], out, x, int(axis))
    return out


def matmul(a: HostView, b: HostView) -> HostView:
    xs, ys = a.shape, b.shape
    nd = max(len(xs), len(ys))
    batch = []
    for i in range(nd - 2):
        ai, bi = i - (nd - len(xs)), i - (nd - len(ys))
        sa = xs[ai] if ai >= 0 else 1
        sb = ys[bi] if bi >= 0 else 1
        batch.append(max(sa, sb))
    out = HostView.empty(a.dtype, batch + [xs[-2], ys[-1]])
    call("matmul", out, a, b)
    return out


def pad(x: HostView, padding, fill_scalar: HostView) -> HostView:
    out_shape = [d + b + a for d, (b, a) in zip(x.shape, padding)]
    out = HostView.empty(x.dtype, out_shape)
    call("pad", out, x, fill_scalar, [b for b, _ in padding])
    return out


def cat(xs, axis: int) -> HostView:
    first = xs[0]
    total = sum(t.shape[axis] for t in xs)
    out_shape = [total if i == axis else d for i, d in enumerate(first.shape)]
    out = HostView.empty(first.dtype, out_shape)
    call("cat", out, list(xs), int(axis))
    return out


def gather(data, indices, axis) -> HostView:
    out = HostView.empty(data.dtype, indices.shape)
    call("gather", out, data, indices, int(axis))
    return out


def scatter(template, indices, updates, axis, mode) -> HostView:
    out = copy(template)
    call("scatter", out, indices, updates, int(axis), {"set": 0, "add": 1}[mode])
    return out


def unfold(x, kernel_size, stride, dilation, padding) -> HostView:
    k = len(kernel_size)
    ld = len(x.shape) - k
    sp = x.shape[ld:]
    # OCaml's `/` truncates toward zero (a kernel wider than the padded extent gives 0 or 1 windows)
    osp = [int(((sp[i] + padding[i][0] + padding[i][1]) - (dilation[i] * (kernel_size[i] - 1) + 1)) / stride[i]) + 1
           for i in range(k)]
    out = HostView.empty(x.dtype, list(x.shape[:ld]) + [int(np.prod(kernel_size)), int(np.prod(osp))])
    call("unfold", out, x, list(kernel_size), list(stride), list(dilation), [v for p in padding for v in p])
    return out


def fold(x, output_size, kernel_size, stride, dilation, padding) -> HostView:
    out = HostView.empty(x.dtype, list(x.shape[:len(x.shape) - 2]) + list(output_size))
    call("fold", out, x, list(output_size), list(kernel_size), list(stride), list(dilation),
         [v for p in padding for v in p])
    return out


def threefry(key, ctr) -> HostView:
    out = HostView.empty("i32", ctr.shape)
    call("threefry", out, key, ctr)
    return out


def sort(x, axis, descending=False) -> HostView:
    out = HostView.empty(x.dtype, x.shape)
    call("sort", out, x, int(axis), bool(descending))
    return out


def argsort(x, axis, descending=False) -> HostView:
    out = HostView.empty("i32", x.shape)
    call("argsort", out, x, int(axis), bool(descending))
    return out


# fft family: the binding owns the output shape (reference: backend_c/nx_backend.ml:520-549)
def fft(x, axes, inverse=False) -> HostView:
    out = HostView.empty(x.dtype, x.shape)
    call("ifft" if inverse else "fft", out, x, [int(a) for a in axes])
    return out


def rfft(x, dtype, axes) -> HostView:
    axes = [int(a) for a in axes]
    shape = list(x.shape)
    shape[axes[-1]] = shape[axes[-1]] // 2 + 1
    out = HostView.empty(dtype, shape)
    call("rfft", out, x, axes)
    return out


def irfft(x, dtype, axes, s=None) -> HostView:
    axes = [int(a) for a in axes]
    shape = list(x.shape)
    shape[axes[-1]] = int(s[-1]) if s is not None else (shape[axes[-1]] - 1) * 2
    out = HostView.empty(dtype, shape)
    call("irfft", out, x, axes, [int(v) for v in s] if s is not None else [])
    return out


# linalg tier 1: the binding owns the output shapes (reference: backend_c/nx_backend.ml:551-640)
def cholesky(x, upper=False) -> HostView:
    out = HostView.empty(x.dtype, x.shape)
    call("cholesky", out, x, bool(upper))
    return out


def triangular_solve(a, b, upper=False, transpose=False, unit_diag=False) -> HostView:
    vector_rhs = len(b.shape) == len(a.shape) - 1
    bm = b.reshape_contig(list(b.shape) + [1]) if vector_rhs else b
    out = HostView.empty(b.dtype, bm.shape)
    call("triangular_solve", out, a, bm, (1 if upper else 0) | (2 if transpose else 0) | (4 if unit_diag else 0))
    return out.reshape_contig(list(b.shape)) if vector_rhs else out


def qr(x, reduced=True):
    m, n = x.shape[-2], x.shape[-1]
    k = min(m, n)
    qs, rs = list(x.shape), list(x.shape)
    if reduced:
        qs[-1], rs[-2] = k, k
    else:
        qs[-1] = m
    q, r = HostView.empty(x.dtype, qs), HostView.empty(x.dtype, rs)
    call("qr", q, r, x, bool(reduced))
    return q, r


# linalg tiers 2-3 (reference: backend_c/nx_backend.ml:627-720): eigenvalues / singular values
# are always f64, eig's outputs always c64
def eigh(x, vectors=True):
    w = HostView.empty("f64", list(x.shape[:-2]) + [x.shape[-1]])
    if not vectors:
        call("eigh", w, x, x, False)
        return w
    v = HostView.empty(x.dtype, x.shape)
    call("eigh", w, v, x, True)
    return w, v


def svd(x, full_matrices=False):
    m, n = x.shape[-2], x.shape[-1]
    k = min(m, n)
    batch = list(x.shape[:-2])
    u = HostView.empty(x.dtype, batch + ([m, m] if full_matrices else [m, k]))
    sv = HostView.empty("f64", batch + [k])
    vt = HostView.empty(x.dtype, batch + ([n, n] if full_matrices else [k, n]))
    call("svd", u, sv, vt, x)
    return u, sv, vt


def eig(x, vectors=True):
    w = HostView.empty("c64", list(x.shape[:-2]) + [x.shape[-1]])
    if not vectors:
        call("eig", w, w, x, False)
        return w
    v = HostView.empty("c64", x.shape)
    call("eig", w, v, x, True)
    return w, v


__all__ = [n for n in dir() if not n.startswith("_")]
