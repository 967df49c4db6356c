"""oracle/nxo.py -- TEST INFRASTRUCTURE ONLY.

ctypes front end of oracle/libnxo.so, the C restatement of the reference's hot
path (oracle/nxo.c). Same function surface as oracle/ref.py (which drives the
reference's own compiled C), so a test can use either as the checker. The veneer
logic (output shapes, sorted axes, empty-axis pre-checks) restates
packages/nx/lib/backend_c/nx_backend.ml:170-500.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module. The product path never does.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .hostview import HostView

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnxo.so")

UNARY = ("neg recip abs sign sqrt exp log sin cos tan asin acos atan sinh cosh tanh "
         "trunc ceil floor round erf").split()
BINARY = "add sub mul idiv fdiv mod max min pow atan2 xor or and shl shr".split()
CMP = "cmpeq cmpne cmplt cmple".split()
REDUCE = {"sum": 0, "prod": 1, "max": 2, "min": 3}


class _T(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("dtype", ctypes.c_int32), ("ndim", ctypes.c_int32),
                ("shape", ctypes.c_int64 * 32), ("strides", ctypes.c_int64 * 32), ("offset", ctypes.c_int64)]


class RefError(Exception):
    def __init__(self, kind, msg):
        super().__init__(f"{kind}: {msg}")
        self.kind = kind
        self.msg = msg


_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(LIB_PATH)
        for n in ("nxo_map1 nxo_map2 nxo_cmp nxo_where nxo_copy nxo_cast nxo_reduce nxo_argreduce nxo_scan "
                  "nxo_matmul nxo_pad nxo_cat nxo_gather nxo_scatter nxo_threefry nxo_fill nxo_sort nxo_unfold nxo_fold").split():
            getattr(_lib, n).restype = ctypes.c_char_p
        _lib.nxo_status_is_invalid_argument.argtypes = [ctypes.c_char_p]
    return _lib


def _d(hv: HostView) -> _T:
    from .hostview import tag
    t = _T()
    t.data = hv.storage.ctypes.data if hv.storage.size else 0
    t.dtype = tag(hv.dtype)
    t.ndim = len(hv.shape)
    for i, (s, st) in enumerate(zip(hv.shape, hv.strides)):
        t.shape[i] = s
        t.strides[i] = st
    t.offset = hv.offset
    return t


def _chk(op, status):
    if status:
        kind = "Invalid_argument" if lib().nxo_status_is_invalid_argument(status) else "Failure"
        raise RefError(kind, f"{op}: {status.decode()}")


def unary(op, x):
    out = HostView.empty(x.dtype, x.shape)
    _chk(op, lib().nxo_map1(UNARY.index(op), ctypes.byref(_d(out)), ctypes.byref(_d(x))))
    return out


def binary(op, x, y):
    out = HostView.empty(x.dtype, x.shape)
    _chk(op, lib().nxo_map2(BINARY.index(op), ctypes.byref(_d(out)), ctypes.byref(_d(x)), ctypes.byref(_d(y))))
    return out


def compare(op, x, y):
    out = HostView.empty("bool", x.shape)
    _chk(op, lib().nxo_cmp(CMP.index(op), ctypes.byref(_d(out)), ctypes.byref(_d(x)), ctypes.byref(_d(y))))
    return out


def where(c, a, b):
    out = HostView.empty(a.dtype, a.shape)
    _chk("where", lib().nxo_where(ctypes.byref(_d(out)), ctypes.byref(_d(c)), ctypes.byref(_d(a)), ctypes.byref(_d(b))))
    return out


def cast(x, dtype):
    out = HostView.empty(dtype, x.shape)
    _chk("cast", lib().nxo_cast(ctypes.byref(_d(out)), ctypes.byref(_d(x))))
    return out


def copy(x):
    out = HostView.empty(x.dtype, x.shape)
    _chk("copy", lib().nxo_copy(ctypes.byref(_d(out)), ctypes.byref(_d(x))))
    return out


def assign(dst, src):
    _chk("copy", lib().nxo_copy(ctypes.byref(_d(dst)), ctypes.byref(_d(src))))


def reduce(op, x, axes):
    axes = sorted(int(a) for a in axes)
    if op in ("max", "min"):
        for ax in axes:
            if x.shape[ax] == 0:
                raise RefError("Invalid_argument", f"reduce_{op}: reduction over an empty axis has no identity")
    out = HostView.empty(x.dtype, [d for i, d in enumerate(x.shape) if i not in axes])
    ax = (ctypes.c_int * max(len(axes), 1))(*axes)
    _chk("reduce_" + op, lib().nxo_reduce(REDUCE[op], ctypes.byref(_d(out)), ctypes.byref(_d(x)), ax, len(axes)))
    return out


def argreduce(op, x, axis, keepdims=False):
    if x.shape[axis] == 0:
        raise RefError("Invalid_argument", f"{op}: argument reduction over an empty axis")
    shape = [1 if i == axis else d for i, d in enumerate(x.shape)] if keepdims else \
        [d for i, d in enumerate(x.shape) if i != axis]
    out = HostView.empty("i32", shape)
    _chk(op, lib().nxo_argreduce(1 if op == "argmax" else 0, ctypes.byref(_d(out)), ctypes.byref(_d(x)), int(axis)))
    return out


def scan(op, x, axis):
    out = HostView.empty(x.dtype, x.shape)
    name = {"sum": "cumsum", "prod": "cumprod", "max": "cummax", "min": "cummin"}[op]
    _chk(name, lib().nxo_scan(REDUCE[op], ctypes.byref(_d(out)), ctypes.byref(_d(x)), int(axis)))
    return out


def matmul(a, b):
    xs, ys = a.shape, b.shape
    nd = max(len(xs), len(ys))
    batch = []
    for i in range(nd - 2):
        ai, bi = i - (nd - len(xs)), i - (nd - len(ys))
        batch.append(max(xs[ai] if ai >= 0 else 1, ys[bi] if bi >= 0 else 1))
    out = HostView.empty(a.dtype, batch + [xs[-2], ys[-1]])
    _chk("matmul", lib().nxo_matmul(ctypes.byref(_d(out)), ctypes.byref(_d(a)), ctypes.byref(_d(b))))
    return out


def pad(x, padding, fill_scalar: HostView):
    out = HostView.empty(x.dtype, [d + b + a for d, (b, a) in zip(x.shape, padding)])
    before = (ctypes.c_int64 * max(len(padding), 1))(*[b for b, _ in padding])
    fill = fill_scalar.numpy().reshape(-1)[:1].copy()
    _chk("pad", lib().nxo_pad(ctypes.byref(_d(out)), ctypes.byref(_d(x)), ctypes.c_void_p(fill.ctypes.data), before))
    return out


def cat(xs, axis):
    first = xs[0]
    total = sum(t.shape[axis] for t in xs)
    out = HostView.empty(first.dtype, [total if i == axis else d for i, d in enumerate(first.shape)])
    ds = [_d(x) for x in xs]
    arr = (ctypes.POINTER(_T) * len(ds))(*[ctypes.pointer(d) for d in ds])
    _chk("cat", lib().nxo_cat(ctypes.byref(_d(out)), arr, len(ds), int(axis)))
    return out


def gather(data, indices, axis):
    out = HostView.empty(data.dtype, indices.shape)
    _chk("gather", lib().nxo_gather(ctypes.byref(_d(out)), ctypes.byref(_d(data)), ctypes.byref(_d(indices)), int(axis)))
    return out


def scatter(template, indices, updates, axis, mode):
    out = copy(template)
    _chk("scatter", lib().nxo_scatter(ctypes.byref(_d(out)), ctypes.byref(_d(indices)), ctypes.byref(_d(updates)),
                                      int(axis), {"set": 0, "add": 1}[mode]))
    return out


def _i64(xs):
    return (ctypes.c_int64 * max(len(xs), 1))(*[int(v) for v in xs])


def unfold(x, kernel_size, stride, dilation, padding):
    k = len(kernel_size)
    ld = len(x.shape) - k
    sp = x.shape[ld:]
    # OCaml's `/` truncates toward zero (a kernel wider than the padded extent gives 0 or 1 windows)
    osp = [int(((sp[i] + padding[i][0] + padding[i][1]) - (dilation[i] * (kernel_size[i] - 1) + 1)) / stride[i]) + 1
           for i in range(k)]
    out = HostView.empty(x.dtype, list(x.shape[:ld]) + [int(np.prod(kernel_size)), int(np.prod(osp))])
    _chk("unfold", lib().nxo_unfold(ctypes.byref(_d(out)), ctypes.byref(_d(x)), k, _i64(kernel_size), _i64(stride),
                                    _i64(dilation), _i64([v for p in padding for v in p])))
    return out


def fold(x, output_size, kernel_size, stride, dilation, padding):
    out = HostView.empty(x.dtype, list(x.shape[:len(x.shape) - 2]) + list(output_size))
    _chk("fold", lib().nxo_fold(ctypes.byref(_d(out)), ctypes.byref(_d(x)), len(kernel_size), _i64(output_size),
                                _i64(kernel_size), _i64(stride), _i64(dilation), _i64([v for p in padding for v in p])))
    return out


def threefry(key, ctr):
    out = HostView.empty("i32", ctr.shape)
    _chk("threefry", lib().nxo_threefry(ctypes.byref(_d(out)), ctypes.byref(_d(key)), ctypes.byref(_d(ctr))))
    return out


def sort(x, axis, descending=False):
    out = HostView.empty(x.dtype, x.shape)
    _chk("sort", lib().nxo_sort(0, ctypes.byref(_d(out)), ctypes.byref(_d(x)), int(axis), 1 if descending else 0))
    return out


def argsort(x, axis, descending=False):
    out = HostView.empty("i32", x.shape)
    _chk("argsort", lib().nxo_sort(1, ctypes.byref(_d(out)), ctypes.byref(_d(x)), int(axis), 1 if descending else 0))
    return out


# ---- fft family -----------------------------------------------------------------------------
# Restated from the DEFINITION the reference pins (nx_c_fft.c:38-41, 940-1143), not from its
# mixed-radix algorithm: an unnormalised direct DFT in double precision, X[k] = sum_j x[j] *
# exp(sign*2*pi*i*(j*k mod n)/n), one axis at a time in the reference's pass order, rounding to the
# output type after every pass exactly where the reference stores (fft/rfft write `out` per axis;
# irfft keeps a c64 temporary for the non-last axes). O(n^2): small cases only.
def _dft_matrix(n, sign):
    jk = (np.arange(n, dtype=np.int64)[:, None] * np.arange(n, dtype=np.int64)[None, :]) % max(n, 1)
    ang = 2.0 * jk.astype(np.float64) / max(n, 1)
    return np.cos(np.pi * ang) + 1j * sign * np.sin(np.pi * ang)


def _dft_axis(a, axis, sign):
    n = a.shape[axis]
    if n == 0:
        return a
    w = _dft_matrix(n, sign)
    return np.moveaxis(np.tensordot(w, np.moveaxis(a, axis, 0), axes=(1, 0)), 0, axis)


def _cplx(dtype):
    if dtype not in ("c32", "c64"):
        raise RefError("Failure", "unsupported bigarray kind")
    return np.complex64 if dtype == "c32" else np.complex128


def _real(dtype):
    if dtype not in ("f32", "f64"):
        raise RefError("Failure", "unsupported bigarray kind")
    return np.float32 if dtype == "f32" else np.float64


def fft(x, axes, inverse=False):
    ct = _cplx(x.dtype)
    a = x.numpy()
    for ax in axes:
        if ax < 0 or ax >= a.ndim:
            raise RefError("Invalid_argument", "axis out of range")
        a = _dft_axis(a.astype(np.complex128), int(ax), 1 if inverse else -1).astype(ct)
    return HostView.from_array(a, x.dtype)


def rfft(x, dtype, axes):
    _real(x.dtype)
    ct = _cplx(dtype)
    axes = [int(v) for v in axes]
    a = x.numpy().astype(np.complex128)
    last = axes[-1]
    half = a.shape[last] // 2 + 1
    a = np.take(_dft_axis(a, last, -1), np.arange(half), axis=last).astype(ct)
    for ax in axes[:-1]:
        a = _dft_axis(a.astype(np.complex128), ax, -1).astype(ct)
    return HostView.from_array(a, dtype)


def irfft(x, dtype, axes, s=None):
    _cplx(x.dtype)
    rt = _real(dtype)
    axes = [int(v) for v in axes]
    a = x.numpy().astype(np.complex128)
    for ax in axes[:-1]:
        a = _dft_axis(a, ax, 1)
    last = axes[-1]
    in_half = a.shape[last]
    size = int(s[-1]) if s is not None else 2 * (in_half - 1)
    half = min(in_half, size // 2 + 1)
    g = np.moveaxis(a, last, -1)
    f = np.zeros(g.shape[:-1] + (size,), dtype=np.complex128)
    f[..., :half] = g[..., :half]
    for k in range(1, half):
        if size - k != k:
            f[..., size - k] = np.conj(g[..., k])
    out = _dft_axis(f, f.ndim - 1, 1).real.astype(rt)
    return HostView.from_array(np.moveaxis(out, -1, last), dtype)
