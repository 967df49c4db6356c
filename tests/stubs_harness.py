"""Drives packages/nx-cuda/lib/nx_cuda_stubs.c -- the C half of the OCaml binding -- without OCaml.

tests/ocaml_rt/libnxcuda_stubs_test.so is that file, unmodified, compiled against the OCaml C API
declarations (tests/ocaml_api, oracle/caml_shim) plus tests/ocaml_rt/runtime_shim.c (custom-block
allocation, the two raisers as longjmp, caml_stat_*). This module fabricates the OCaml values the
stubs take -- immediates, int arrays, bigarrays, custom blocks, and the tensor RECORD in the field
order packages/nx-cuda/lib/nx_backend.ml declares (buffer; shape; strides; offset; tag; context;
dtype; elems) -- and restates what that veneer does around each `external` (allocate the output,
compute its shape, pass (out, inputs...)), so that every stub runs on the GPU against the oracle.
What stays unverified is then only the OCaml text of nx_backend.ml itself.

TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess
import weakref

import numpy as np

from oracle.hostview import HostView, c_strides, np_storage, numel, tag as dtype_tag
from oracle.ref import _KIND
from raven_b200._lib import Failure, InvalidArgument
from raven_b200.view import View

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "tests", "ocaml_rt", "libnxcuda_stubs_test.so")
_lib = None
U64 = 0xFFFFFFFFFFFFFFFF


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "ocaml_rt")], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.nxstub_invoke.restype = ctypes.c_int
        _lib.nxstub_invoke.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        _lib.nxstub_last_message.restype = ctypes.c_char_p
        _lib.nxstub_release.argtypes = [ctypes.c_uint64]
        _lib.nxstub_live.restype = ctypes.c_long
        _lib.nxstub_custom_identifier.restype = ctypes.c_char_p
        _lib.nxstub_custom_identifier.argtypes = [ctypes.c_uint64]
    return _lib


def val_int(n):
    return ((int(n) << 1) | 1) & U64


class _Args:
    """OCaml blocks built for one call; kept alive until it returns."""

    def __init__(self):
        self.keep = []

    def block(self, fields):
        words = (ctypes.c_uint64 * (len(fields) + 1))()
        words[0] = len(fields) << 10
        for i, f in enumerate(fields):
            words[i + 1] = f & U64
        self.keep.append(words)
        return ctypes.addressof(words) + 8

    def ints(self, xs):
        return self.block([val_int(x) for x in xs])

    def bigarray(self, arr: np.ndarray, dtype: str):
        kind, ext = _KIND[dtype]
        self.keep.append(arr)
        return self.block([0, arr.ctypes.data if arr.size else 0, 1, kind | (ext << 16), 0, arr.size])


def invoke(name, nargs, vals, keep=None):
    L = lib()
    fn = ctypes.cast(getattr(L, name), ctypes.c_void_p)
    arr = (ctypes.c_uint64 * max(len(vals), 1))(*[v & U64 for v in vals])
    out = ctypes.c_uint64(0)
    rc = L.nxstub_invoke(fn, nargs, ctypes.cast(arr, ctypes.c_void_p), ctypes.byref(out))
    del keep
    if rc == 1:
        raise Failure(L.nxstub_last_message().decode())
    if rc == 2:
        raise InvalidArgument(L.nxstub_last_message().decode())
    if rc != 0:
        raise RuntimeError(f"nxstub_invoke: bad arity for {name}")
    return out.value


class Custom:
    """A custom block returned by a stub (context or device buffer); released -- finalizer run --
    when the wrapper is collected, as the GC would."""

    def __init__(self, value):
        self.value = value
        weakref.finalize(self, lib().nxstub_release, value)


class ST:
    """The veneer's tensor record."""

    def __init__(self, buf, shape, strides, offset, dtype, ctx, elems):
        self.buf, self.shape, self.strides, self.offset = buf, tuple(shape), tuple(strides), int(offset)
        self.dtype, self.ctx, self.elems = dtype, ctx, int(elems)

    def record(self, a: _Args):
        # FIELD ORDER IS ABI (nx_backend.ml:20-29): buffer shape strides offset tag context dtype elems
        return a.block([self.buf.value, a.ints(self.shape), a.ints(self.strides), val_int(self.offset),
                        val_int(dtype_tag(self.dtype)), self.ctx.value, val_int(0), val_int(self.elems)])


UNARY = ("neg recip abs sign sqrt exp log sin cos tan asin acos atan sinh cosh tanh trunc ceil floor round erf").split()
BINARY = "add sub mul idiv fdiv mod max min pow atan2 xor or and shl shr".split()
CMP = "cmpeq cmpne cmplt cmple".split()
REDUCE = {"sum": 0, "prod": 1, "max": 2, "min": 3}


class StubBackend:
    """nx_backend.ml, restated over the stubs."""

    def __init__(self):
        self.ctx = Custom(invoke("nx_cuda_ctx_create", 1, [val_int(0)]))
        assert lib().nxstub_custom_identifier(self.ctx.value) == b"nx_cuda.ctx"

    # ---- creation / transfer ----
    def _esize(self, dtype):
        return np.dtype(np_storage(dtype)).itemsize

    def create(self, dtype, shape):
        n = numel(shape)
        nbytes = (n + 1) // 2 if dtype in ("i4", "u4") else n * self._esize(dtype)
        buf = Custom(invoke("nx_cuda_alloc", 2, [self.ctx.value, val_int(max(16, nbytes))]))
        return ST(buf, shape, c_strides(shape), 0, dtype, self.ctx, n)

    def from_host(self, storage: np.ndarray, dtype):
        a = _Args()
        storage = np.ascontiguousarray(storage)
        buf = Custom(invoke("nx_cuda_of_host", 2, [self.ctx.value, a.bigarray(storage, dtype)], a))
        n = storage.size * 2 if dtype in ("i4", "u4") else storage.size
        return ST(buf, [n], [1], 0, dtype, self.ctx, n)

    def upload(self, hv: HostView):
        base = self.from_host(hv.storage, hv.dtype)
        return ST(base.buf, hv.shape, hv.strides, hv.offset, hv.dtype, self.ctx, base.elems)

    def to_host(self, t: ST) -> np.ndarray:
        n = (t.elems + 1) // 2 if t.dtype in ("i4", "u4") else t.elems
        host = np.zeros(n, dtype=np_storage(t.dtype))
        a = _Args()
        invoke("nx_cuda_to_host", 3, [self.ctx.value, t.buf.value, a.bigarray(host, t.dtype)], a)
        return host

    def download(self, t: ST) -> np.ndarray:
        c = t if self.is_c_contiguous(t) else self.copy(t)
        flat = self.to_host(c)
        return flat[c.offset:c.offset + numel(c.shape)].reshape(c.shape)

    def full(self, dtype, shape, storage_scalar: np.ndarray):
        t = self.create(dtype, shape)
        a = _Args()
        invoke("nx_cuda_fill", 2, [t.record(a), a.bigarray(np.ascontiguousarray(storage_scalar).reshape(1), dtype)], a)
        return t

    # ---- movement ----
    def _view(self, t, v):
        return ST(t.buf, v.shape, v.strides, v.offset, t.dtype, t.ctx, t.elems)

    def is_c_contiguous(self, t):
        return t.offset == 0 and tuple(t.strides) == tuple(c_strides(t.shape))

    def permute(self, t, axes):
        return self._view(t, View(t.shape, t.strides, t.offset).permute(axes))

    # ---- map family ----
    def _call(self, name, nargs, build):
        a = _Args()
        return invoke(name, nargs, build(a), a)

    def unary(self, op, x):
        out = self.create(x.dtype, x.shape)
        self._call("nx_cuda_map1", 3, lambda a: [val_int(UNARY.index(op)), out.record(a), x.record(a)])
        return out

    def binary(self, op, x, y):
        out = self.create(x.dtype, x.shape)
        self._call("nx_cuda_map2", 4, lambda a: [val_int(BINARY.index(op)), out.record(a), x.record(a), y.record(a)])
        return out

    def raw_map1(self, code, x):
        out = self.create(x.dtype, x.shape)
        self._call("nx_cuda_map1", 3, lambda a: [val_int(code), out.record(a), x.record(a)])
        return out

    def compare(self, op, x, y):
        out = self.create("bool", x.shape)
        self._call("nx_cuda_cmp", 4, lambda a: [val_int(CMP.index(op)), out.record(a), x.record(a), y.record(a)])
        return out

    def where(self, c, x, y):
        out = self.create(x.dtype, x.shape)
        self._call("nx_cuda_where", 4, lambda a: [out.record(a), c.record(a), x.record(a), y.record(a)])
        return out

    def cast(self, x, dtype):
        out = self.create(dtype, x.shape)
        self._call("nx_cuda_cast", 2, lambda a: [out.record(a), x.record(a)])
        return out

    def copy(self, x):
        out = self.create(x.dtype, x.shape)
        self._call("nx_cuda_copy", 2, lambda a: [out.record(a), x.record(a)])
        return out

    def assign(self, dst, src):
        self._call("nx_cuda_copy", 2, lambda a: [dst.record(a), src.record(a)])

    # ---- fold family ----
    def reduce(self, op, x, axes):
        axes = sorted(axes)
        out = self.create(x.dtype, [d for i, d in enumerate(x.shape) if i not in axes])
        self._call("nx_cuda_reduce", 4, lambda a: [val_int(REDUCE[op]), out.record(a), x.record(a), a.ints(axes)])
        return out

    def argreduce(self, op, x, axis, keepdims=False):
        shp = [1 if i == axis else d for i, d in enumerate(x.shape)] if keepdims else \
            [d for i, d in enumerate(x.shape) if i != axis]
        out = self.create("i32", shp)
        self._call("nx_cuda_argreduce", 4, lambda a: [val_int(1 if op == "argmax" else 0), out.record(a), x.record(a),
                                                      val_int(axis)])
        return out

    def scan(self, op, x, axis):
        out = self.create(x.dtype, x.shape)
        self._call("nx_cuda_scan", 4, lambda a: [val_int(REDUCE[op]), out.record(a), x.record(a), val_int(axis)])
        return out

    def matmul(self, x, y):
        nd = max(len(x.shape), len(y.shape))
        batch = []
        for i in range(nd - 2):
            ai, bi = i - (nd - len(x.shape)), i - (nd - len(y.shape))
            batch.append(max(x.shape[ai] if ai >= 0 else 1, y.shape[bi] if bi >= 0 else 1))
        out = self.create(x.dtype, batch + [x.shape[-2], y.shape[-1]])
        self._call("nx_cuda_matmul", 3, lambda a: [out.record(a), x.record(a), y.record(a)])
        return out

    # ---- move family ----
    def pad(self, x, padding, storage_scalar):
        out = self.create(x.dtype, [d + b + e for d, (b, e) in zip(x.shape, padding)])
        self._call("nx_cuda_pad", 4, lambda a: [out.record(a), x.record(a),
                                                a.bigarray(np.ascontiguousarray(storage_scalar).reshape(1), x.dtype),
                                                a.ints([b for b, _ in padding])])
        return out

    def cat(self, xs, axis):
        shp = list(xs[0].shape)
        shp[axis] = sum(t.shape[axis] for t in xs)
        out = self.create(xs[0].dtype, shp)
        self._call("nx_cuda_cat", 3, lambda a: [out.record(a), a.block([t.record(a) for t in xs]), val_int(axis)])
        return out

    def gather(self, data, idx, axis):
        out = self.create(data.dtype, idx.shape)
        self._call("nx_cuda_gather", 4, lambda a: [out.record(a), data.record(a), idx.record(a), val_int(axis)])
        return out

    def scatter(self, template, idx, upd, axis, mode):
        out = self.copy(template)
        self._call("nx_cuda_scatter", 5, lambda a: [out.record(a), idx.record(a), upd.record(a), val_int(axis),
                                                    val_int({"set": 0, "add": 1}[mode])])
        return out

    def threefry(self, key, ctr):
        out = self.create("i32", ctr.shape)
        self._call("nx_cuda_threefry", 3, lambda a: [out.record(a), key.record(a), ctr.record(a)])
        return out

    def sort(self, x, axis, descending=False, arg=False):
        out = self.create("i32" if arg else x.dtype, x.shape)
        self._call("nx_cuda_sort", 5, lambda a: [val_int(1 if arg else 0), out.record(a), x.record(a), val_int(axis),
                                                 val_int(1 if descending else 0)])
        return out

    def unfold(self, x, kernel, stride, dilation, padding, bytecode=False):
        k = len(kernel)
        lead, spatial = list(x.shape[:len(x.shape) - k]), x.shape[len(x.shape) - k:]
        outsp = [int(((spatial[i] + padding[i][0] + padding[i][1]) - (dilation[i] * (kernel[i] - 1) + 1)) / stride[i]) + 1
                 for i in range(k)]
        out = self.create(x.dtype, lead + [int(np.prod(kernel)), int(np.prod(outsp))])
        flat = [v for pr in padding for v in pr]
        build = lambda a: [out.record(a), x.record(a), a.ints(kernel), a.ints(stride), a.ints(dilation), a.ints(flat)]  # noqa: E731
        if bytecode:   # > 5 arguments: the bytecode entry takes (argv, argn)
            self._call("nx_cuda_unfold_bc", -6, build)
        else:
            self._call("nx_cuda_unfold", 6, build)
        return out

    def cholesky(self, x, upper=False):
        out = self.create(x.dtype, x.shape)
        self._call("nx_cuda_cholesky", 3, lambda a: [out.record(a), x.record(a), val_int(1 if upper else 0)])
        return out

    def fft(self, x, axes, inverse=False):
        out = self.create(x.dtype, x.shape)
        self._call("nx_cuda_fft", 4, lambda a: [val_int(1 if inverse else 0), out.record(a), x.record(a), a.ints(axes)])
        return out

    # ---- step capture ----
    def capture_begin(self):
        invoke("nx_cuda_capture_begin", 1, [self.ctx.value])

    def capture_end(self):
        return Custom(invoke("nx_cuda_capture_end", 1, [self.ctx.value]))

    def graph_launch(self, g):
        invoke("nx_cuda_graph_launch", 2, [self.ctx.value, g.value])

    def sync(self):
        invoke("nx_cuda_sync", 1, [self.ctx.value])
