"""nx-cuda host mirror of the reference's `Nx_backend` (the `nx.backend` seam).

Every function below has the name, argument meaning and error behaviour of the
corresponding entry of `Nx_core.Backend_intf.S` (reference:
packages/nx/lib/core/backend_intf.ml:77-721) as the reference's C-backend veneer
fills it (reference: packages/nx/lib/backend_c/nx_backend.ml:36-500): compute
ops allocate a fresh C-contiguous output and hand the strided operands to the
engine; movement ops are pure view rewrites sharing the buffer; `reduce` drops
the reduced axes; `argmax/argmin` honour keepdims; errors are raised as
"<op>: <msg>" in the class the reference uses (InvalidArgument ~
Invalid_argument, Failure ~ Failure).

The OCaml toolchain is absent from this image, so this Python layer is the
host side above the C ABI; `packages/nx-cuda/` holds the OCaml binding a
maintainer would compile. Everything semantic lives in
libnxcuda.so, so both are plumbing.

Device memory is owned by the engine's stream-ordered caching allocator; a
buffer is returned to it when the last handle sharing it is collected, ordered
after in-flight kernels (cudaFreeAsync on the context stream).
"""
from __future__ import annotations

import builtins as _b
import ctypes

import numpy as np

from . import dtype as _dt
from ._lib import Failure, InvalidArgument, NxcTensor, check, load
from .view import View, c_contiguous_strides, numel

UNARY_OPS = ("neg recip abs sign sqrt exp log sin cos tan asin acos atan sinh cosh tanh "
             "trunc ceil floor round erf").split()
BINARY_OPS = "add sub mul idiv fdiv mod max min pow atan2 xor or and shl shr".split()
CMP_OPS = "cmpeq cmpne cmplt cmple".split()
REDUCE_OPS = {"sum": 0, "prod": 1, "max": 2, "min": 3}


class Context:
    """`create_context : unit -> context` (reference: backend/nx_backend.mli:32-39).
    Device index comes from NX_CUDA_DEVICE / LOCAL_RANK unless given."""

    def __init__(self, device=None, stream=None):
        lib = load()
        p = ctypes.c_void_p()
        if device is None and stream is None:
            st = lib.nxc_ctx_create(ctypes.byref(p))
        else:
            st = lib.nxc_ctx_create_on(-1 if device is None else int(device), stream, ctypes.byref(p))
        check(None, "create_context", st)
        self._p = p
        self._lib = lib
        self._pinned = []   # [lo, hi) address ranges handed out by pinned_empty
        self._capturing = None   # the Capture being recorded, if any

    @property
    def ptr(self):
        return self._p

    def sync(self):
        check(self._p, "sync", self._lib.nxc_sync(self._p))

    def launch_count(self) -> int:
        return int(self._lib.nxc_launch_count(self._p))

    def pinned_empty(self, n: int, dtype) -> np.ndarray:
        """A 1-D numpy array over page-locked host memory (nxc_host_alloc). `from_host` of such an
        array goes through the upload engine and `to_host_async` writes into one; the memory
        lives as long as the array (or any view of it) does."""
        npdt = np.dtype(dtype)
        p = ctypes.c_void_p()
        check(self._p, "host_alloc", self._lib.nxc_host_alloc(self._p, _b.max(n * npdt.itemsize, 1), ctypes.byref(p)))
        nbytes = _b.max(n * npdt.itemsize, 1)
        raw = (ctypes.c_uint8 * nbytes).from_address(p.value)
        arr = np.frombuffer(raw, dtype=npdt, count=n)
        lib, ctxp, ranges, key = self._lib, self._p, self._pinned, (p.value, p.value + nbytes)
        ranges.append(key)
        import weakref

        def _release():
            ranges.remove(key)
            lib.nxc_host_free(ctxp, p)
        weakref.finalize(raw, _release)
        return arr

    def is_pinned(self, address: int) -> bool:
        return any(lo <= address < hi for lo, hi in self._pinned)

    def stream(self) -> int:
        return int(self._lib.nxc_stream(self._p) or 0)

    def device(self) -> int:
        return int(self._lib.nxc_device(self._p))

    def set_matmul_mode(self, mode: str):
        if self._lib.nxc_set_matmul_mode(self._p, mode.encode()) != 0:
            raise InvalidArgument(f"set_matmul_mode: unknown mode {mode!r}")

    def capture(self) -> "Capture":
        """`with ctx.capture() as g: outs = step()` records the ops issued inside the block into a
        CUDA graph instead of running them (nxc_capture_begin / nxc_capture_end); `g.launch()` then
        replays the whole step with one launch. Tensors created inside the block are the replay's
        outputs: they keep their addresses, are rewritten by every `g.launch()`, and stay valid as
        long as `g` does (each holds a reference to it). Tensors that existed before the block are
        read in place by every replay -- keep them alive, refresh them with `assign`."""
        return Capture(self)

    def close(self):
        if self._p:
            self._lib.nxc_ctx_destroy(self._p)
            self._p = None


class Capture:
    """A captured step (see Context.capture)."""

    def __init__(self, ctx: "Context"):
        self.ctx = ctx
        self._g = None
        self.kernels = 0
        self.arena_bytes = 0

    def __enter__(self):
        check(self.ctx.ptr, "capture_begin", self.ctx._lib.nxc_capture_begin(self.ctx.ptr))
        self.ctx._capturing = self
        return self

    def __exit__(self, et, ev, tb):
        self.ctx._capturing = None
        g = ctypes.c_void_p()
        st = self.ctx._lib.nxc_capture_end(self.ctx.ptr, ctypes.byref(g))
        if et is None:
            check(self.ctx.ptr, "capture_end", st)
            self._g = g
            self.kernels = int(self.ctx._lib.nxc_graph_kernels(g))
            self.arena_bytes = int(self.ctx._lib.nxc_graph_arena_bytes(g))
        elif not st and g:
            self.ctx._lib.nxc_graph_destroy(self.ctx.ptr, g)
        return False

    def launch(self):
        if self._g is None:
            raise Failure("graph_launch: nothing was captured")
        check(self.ctx.ptr, "graph_launch", self.ctx._lib.nxc_graph_launch(self.ctx.ptr, self._g))

    def close(self):
        if self._g is not None and self.ctx._p:
            self.ctx._lib.nxc_graph_destroy(self.ctx.ptr, self._g)
        self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def create_context(device=None, stream=None) -> Context:
    return Context(device, stream)


class _Buffer:
    """A device allocation, shared by every handle viewing it."""

    __slots__ = ("ctx", "ptr", "nbytes", "owned", "_host_src", "_graph", "__weakref__")

    def __init__(self, ctx: Context, nbytes: int, ptr=None):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        self._graph = ctx._capturing   # arena memory lives as long as the captured graph does
        if ptr is None:
            p = ctypes.c_void_p()
            check(ctx.ptr, "buffer", ctx._lib.nxc_alloc(ctx.ptr, self.nbytes, ctypes.byref(p)))
            self.ptr = p.value
            self.owned = True
        else:
            self.ptr = int(ptr)
            self.owned = False

    def __del__(self):
        try:
            # (a handle into a captured step's arena is only counted by the engine: the arena lives
            # until the graph is destroyed AND its last handle is gone, whichever comes last)
            if self.owned and self.ptr and self.ctx._p:
                self.ctx._lib.nxc_free(self.ctx.ptr, self.ptr)
        except Exception:
            pass


class Tensor:
    """The handle `('a,'b) t` = {buffer; shape; strides; offset; dtype; context}
    (reference: backend_c/nx_backend.ml:36-43)."""

    __slots__ = ("buffer", "shape", "strides", "offset", "dtype", "context", "_d")

    def __init__(self, buffer, shape, strides, offset, dtype, context):
        self._d = None
        self.buffer = buffer
        self.shape = tuple(int(s) for s in shape)
        self.strides = tuple(int(s) for s in strides)
        self.offset = int(offset)
        self.dtype = dtype
        self.context = context

    def _desc(self) -> NxcTensor:
        # a handle is immutable and the engine only reads descriptors: build once, reuse per op
        if self._d is not None:
            return self._d
        if len(self.shape) > 32:
            raise Failure("ndim exceeds NX_C_MAX_NDIM")
        t = NxcTensor()
        t.data = self.buffer.ptr
        t.dtype = self.dtype.tag
        t.ndim = len(self.shape)
        for i, (s, st) in enumerate(zip(self.shape, self.strides)):
            t.shape[i] = s
            t.strides[i] = st
        t.offset = self.offset
        self._d = t
        return t

    def __repr__(self):
        return f"Tensor({self.dtype.name}, shape={self.shape}, strides={self.strides}, offset={self.offset})"


# ---- accessors (backend_intf.ml:98-116) ---------------------------------------------
def view(t: Tensor) -> View:
    return View(t.shape, t.strides, t.offset)


def dtype(t: Tensor):
    return t.dtype


def context(t: Tensor) -> Context:
    return t.context


def to_host(t: Tensor) -> np.ndarray:
    """Copy the STORAGE out (device backends copy; backend_intf.ml:108-116). The
    result is the flat buffer; index it with the view's offset/strides, as the
    frontend does (frontend.ml:1703-1708). Blocks until the stream drains."""
    n = t.buffer.nbytes
    host = np.empty(n // _b.max(t.dtype.np.itemsize, 1), dtype=t.dtype.np)
    if n:
        check(t.context.ptr, "to_host",
              t.context._lib.nxc_d2h(t.context.ptr, host.ctypes.data, t.buffer.ptr, host.nbytes))
    return host


def to_host_async(t: Tensor, pinned_out: np.ndarray) -> None:
    """Start copying the STORAGE of `t` into `pinned_out` (from Context.pinned_empty) on the
    download engine and return at once: later kernels do not wait for it and the next step's
    uploads overlap it. The data is valid after `ctx.sync()`. `t` may be dropped right away."""
    n = t.buffer.nbytes
    if pinned_out.nbytes < n:
        raise InvalidArgument("to_host_async: destination smaller than the tensor's storage")
    if n:
        check(t.context.ptr, "to_host_async",
              t.context._lib.nxc_d2h_async(t.context.ptr, pinned_out.ctypes.data, t.buffer.ptr, n))


def to_numpy(t: Tensor) -> np.ndarray:
    """Convenience for tests: the logical elements as a C-contiguous array of the
    storage type (raw bits for f16/bf16/fp8)."""
    c = contiguous(t)
    flat = to_host(c)
    n = numel(c.shape)
    return flat[c.offset:c.offset + n].reshape(c.shape)


# ---- creation (backend_intf.ml:120-138) ----------------------------------------------
def _nbytes(dt, n):
    return (n + 1) // 2 if dt.itemsize == 0 else n * dt.itemsize


def _create(ctx: Context, dt, shape) -> Tensor:
    shape = tuple(int(s) for s in shape)
    buf = _Buffer(ctx, _nbytes(dt, numel(shape)))
    return Tensor(buf, shape, c_contiguous_strides(shape), 0, dt, ctx)


def buffer(ctx: Context, dt, shape) -> Tensor:
    return _create(ctx, _dt.of(dt), shape)


def full(ctx: Context, dt, shape, value) -> Tensor:
    """Device-side fill: the value travels as a kernel argument (no H2D copy, no
    sync), so `x *. c`-style scalar operands cost one tiny launch."""
    dt = _dt.of(dt)
    t = _create(ctx, dt, shape)
    scalar = _dt.encode_scalar(dt, value)
    d = t._desc()
    check(ctx.ptr, "full", ctx._lib.nxc_fill(ctx.ptr, ctypes.byref(d), scalar.ctypes.data))
    return t


def from_host(ctx: Context, array: np.ndarray, dt=None) -> Tensor:
    """1-D view over a host buffer; a device backend copies it
    (backend_intf.ml:132-138). `dt` names the Nx dtype when the numpy type is
    ambiguous (uint16 bits of bf16/f16, uint8 bits of fp8/bool)."""
    a = np.ascontiguousarray(array).reshape(-1)
    if dt is None:
        if a.dtype == np.bool_:
            dt, a = _dt.bool_, a.view(np.uint8)
        elif a.dtype == np.float16:
            dt, a = _dt.float16, a.view(np.uint16)
        else:
            dt = next(d for d in _dt.ALL if d.np == a.dtype and d.name not in
                      ("bf16", "f16", "f8e4m3", "f8e5m2", "bool", "i4", "u4"))
    else:
        dt = _dt.of(dt)
        if a.dtype == np.bool_:
            a = a.view(np.uint8)
        if a.dtype == np.float16:
            a = a.view(np.uint16)
        if a.dtype != dt.np:
            raise InvalidArgument(f"from_host: host array is {a.dtype}, dtype {dt.name} stores {dt.np}")
    n = a.size * 2 if dt.itemsize == 0 else a.size
    t = _create(ctx, dt, (n,))
    if a.nbytes:
        check(ctx.ptr, "from_host", ctx._lib.nxc_h2d(ctx.ptr, t.buffer.ptr, a.ctypes.data, a.nbytes))
        if ctx.is_pinned(a.ctypes.data):
            # page-locked source (Context.pinned_empty): the upload engine reads it asynchronously.
            # The tensor keeps the array alive; the caller must not rewrite it before ctx.sync().
            t.buffer._host_src = a
        else:
            ctx.sync()  # the host array may be freed or rewritten by the caller right away
    return t


# ---- movement: pure view rewrites (backend_c/nx_backend.ml:77-84) ----------------------
def _of_view(t: Tensor, v: View) -> Tensor:
    return Tensor(t.buffer, v.shape, v.strides, v.offset, t.dtype, t.context)


def expand(t, shape):
    return _of_view(t, view(t).expand(shape))


def reshape(t, shape):
    return _of_view(t, view(t).reshape(shape))


def permute(t, axes):
    return _of_view(t, view(t).permute(axes))


def shrink(t, bounds):
    return _of_view(t, view(t).shrink(bounds))


def flip(t, flags):
    return _of_view(t, view(t).flip(flags))


def is_c_contiguous(t: Tensor) -> bool:
    return view(t).is_c_contiguous() and t.offset == 0


# ---- map family (backend_c/nx_backend.ml:170-234) ----------------------------------------
def _call(ctx, op, fn, *args):
    check(ctx.ptr, op, fn(ctx.ptr, *args))


def _unary(opname):
    code = UNARY_OPS.index(opname)

    def f(x: Tensor) -> Tensor:
        out = _create(x.context, x.dtype, x.shape)
        do, dx = out._desc(), x._desc()
        _call(x.context, opname, x.context._lib.nxc_map1, code, ctypes.byref(do), ctypes.byref(dx))
        return out

    f.__name__ = opname
    return f


def _binary(opname):
    code = BINARY_OPS.index(opname)

    def f(x: Tensor, y: Tensor) -> Tensor:
        out = _create(x.context, x.dtype, x.shape)
        do, dx, dy = out._desc(), x._desc(), y._desc()
        _call(x.context, opname, x.context._lib.nxc_map2, code, ctypes.byref(do), ctypes.byref(dx),
              ctypes.byref(dy))
        return out

    f.__name__ = opname
    return f


def _compare(opname):
    code = CMP_OPS.index(opname)

    def f(x: Tensor, y: Tensor) -> Tensor:
        out = _create(x.context, _dt.bool_, x.shape)
        do, dx, dy = out._desc(), x._desc(), y._desc()
        _call(x.context, opname, x.context._lib.nxc_cmp, code, ctypes.byref(do), ctypes.byref(dx),
              ctypes.byref(dy))
        return out

    f.__name__ = opname
    return f


neg, recip, abs, sign, sqrt, exp, log, sin, cos, tan = (_unary(n) for n in UNARY_OPS[:10])  # noqa: A001
asin, acos, atan, sinh, cosh, tanh, trunc, ceil, floor, round, erf = (_unary(n) for n in UNARY_OPS[10:])  # noqa: A001
add, sub, mul, idiv, fdiv = (_binary(n) for n in BINARY_OPS[:5])
mod_ = _binary("mod")
max, min, pow, atan2, xor = (_binary(n) for n in ("max", "min", "pow", "atan2", "xor"))  # noqa: A001
or_ = _binary("or")
and_ = _binary("and")
shl, shr = _binary("shl"), _binary("shr")
cmpeq, cmpne, cmplt, cmple = (_compare(n) for n in CMP_OPS)


def where(cond: Tensor, if_true: Tensor, if_false: Tensor) -> Tensor:
    ctx = if_true.context
    out = _create(ctx, if_true.dtype, if_true.shape)
    do, dc, da, db = out._desc(), cond._desc(), if_true._desc(), if_false._desc()
    _call(ctx, "where", ctx._lib.nxc_where, ctypes.byref(do), ctypes.byref(dc), ctypes.byref(da),
          ctypes.byref(db))
    return out


def cast(x: Tensor, dtype) -> Tensor:  # noqa: A002  (`cast ~dtype x`)
    dt = _dt.of(dtype)
    out = _create(x.context, dt, x.shape)
    do, dx = out._desc(), x._desc()
    _call(x.context, "cast", x.context._lib.nxc_cast, ctypes.byref(do), ctypes.byref(dx))
    return out


# ---- move family (backend_c/nx_backend.ml:349-357) -----------------------------------------
def copy(x: Tensor) -> Tensor:
    out = _create(x.context, x.dtype, x.shape)
    do, dx = out._desc(), x._desc()
    _call(x.context, "copy", x.context._lib.nxc_copy, ctypes.byref(do), ctypes.byref(dx))
    return out


def contiguous(x: Tensor) -> Tensor:
    return x if is_c_contiguous(x) else copy(x)


def assign(dst: Tensor, src: Tensor) -> None:
    dd, ds = dst._desc(), src._desc()
    _call(dst.context, "copy", dst.context._lib.nxc_copy, ctypes.byref(dd), ctypes.byref(ds))


def pad(x: Tensor, padding, fill_value) -> Tensor:
    dt = x.dtype
    out_shape = [d + b + a for d, (b, a) in zip(x.shape, padding)]
    out = _create(x.context, dt, out_shape)
    scalar = _dt.encode_scalar(dt, fill_value)
    before = (ctypes.c_int64 * _b.max(len(padding), 1))(*[b for b, _ in padding])
    do, dx = out._desc(), x._desc()
    _call(x.context, "pad", x.context._lib.nxc_pad, ctypes.byref(do), ctypes.byref(dx),
          scalar.ctypes.data, before)
    return out


def cat(tensors, axis: int) -> Tensor:
    tensors = list(tensors)
    if not tensors:
        raise InvalidArgument("cat: empty tensor list")
    first = tensors[0]
    nd = len(first.shape)
    if axis < 0:
        axis += nd
    total = sum(t.shape[axis] for t in tensors)
    out_shape = [total if i == axis else d for i, d in enumerate(first.shape)]
    out = _create(first.context, first.dtype, out_shape)
    descs = [t._desc() for t in tensors]
    arr = (ctypes.POINTER(NxcTensor) * len(descs))(*[ctypes.pointer(d) for d in descs])
    do = out._desc()
    _call(first.context, "cat", first.context._lib.nxc_cat, ctypes.byref(do), arr, len(descs), axis)
    return out


def gather(data: Tensor, indices: Tensor, axis: int, trusted: bool = False) -> Tensor:
    """`trusted`: the indices come from the backend's own argmax / argmin / argsort, so the
    out-of-range flag is not read back and the call does not drain the stream."""
    out = _create(data.context, data.dtype, indices.shape)
    do, dd, di = out._desc(), data._desc(), indices._desc()
    fn = data.context._lib.nxc_gather_trusted if trusted else data.context._lib.nxc_gather
    _call(data.context, "gather", fn, ctypes.byref(do), ctypes.byref(dd), ctypes.byref(di), axis)
    return out


def scatter(template: Tensor, indices: Tensor, updates: Tensor, axis: int, mode="set",
            unique_indices=False) -> Tensor:
    out = copy(template)
    do, di, du = out._desc(), indices._desc(), updates._desc()
    _call(template.context, "scatter", template.context._lib.nxc_scatter, ctypes.byref(do),
          ctypes.byref(di), ctypes.byref(du), axis, {"set": 0, "add": 1}[mode])
    return out


def threefry(key: Tensor, counter: Tensor) -> Tensor:
    out = _create(counter.context, _dt.int32, counter.shape)
    do, dk, dc = out._desc(), key._desc(), counter._desc()
    _call(counter.context, "threefry", counter.context._lib.nxc_threefry, ctypes.byref(do),
          ctypes.byref(dk), ctypes.byref(dc))
    return out


# ---- fold family (backend_c/nx_backend.ml:266-319) -----------------------------------------
def reduce_output_shape(shape, axes, keepdims):
    if keepdims:
        return [1 if i in axes else d for i, d in enumerate(shape)]
    return [d for i, d in enumerate(shape) if i not in axes]


def reduce(x: Tensor, op: str, axes) -> Tensor:  # noqa: A001  (`reduce ~op ~axes x`)
    axes = sorted(int(a) for a in axes)
    if op in ("max", "min"):
        for ax in axes:
            if x.shape[ax] == 0:
                raise InvalidArgument(f"reduce_{op}: reduction over an empty axis has no identity")
    out = _create(x.context, x.dtype, reduce_output_shape(x.shape, axes, False))
    do, dx = out._desc(), x._desc()
    ax = (ctypes.c_int * _b.max(len(axes), 1))(*axes)
    _call(x.context, "reduce_" + op, x.context._lib.nxc_reduce, REDUCE_OPS[op], ctypes.byref(do),
          ctypes.byref(dx), ax, len(axes))
    return out


def _argreduce(opname, is_max):
    def f(x: Tensor, axis: int, keepdims: bool = False) -> Tensor:
        if x.shape[axis] == 0:
            raise InvalidArgument(f"{opname}: argument reduction over an empty axis")
        out = _create(x.context, _dt.int32, reduce_output_shape(x.shape, [axis], keepdims))
        do, dx = out._desc(), x._desc()
        _call(x.context, opname, x.context._lib.nxc_argreduce, is_max, ctypes.byref(do), ctypes.byref(dx),
              int(axis))
        return out

    f.__name__ = opname
    return f


argmax = _argreduce("argmax", 1)
argmin = _argreduce("argmin", 0)


def associative_scan(x: Tensor, axis: int, op: str) -> Tensor:
    out = _create(x.context, x.dtype, x.shape)
    do, dx = out._desc(), x._desc()
    name = {"sum": "cumsum", "prod": "cumprod", "max": "cummax", "min": "cummin"}[op]
    _call(x.context, name, x.context._lib.nxc_scan, REDUCE_OPS[op], ctypes.byref(do), ctypes.byref(dx),
          int(axis))
    return out


# ---- window ops (backend_c/nx_backend.ml:418-464) ------------------------------------------
def _i64(xs):
    return (ctypes.c_int64 * _b.max(len(xs), 1))(*[int(v) for v in xs])


def unfold(x: Tensor, kernel_size, stride, dilation, padding) -> Tensor:
    k = len(kernel_size)
    lead_nd = len(x.shape) - k
    leading, spatial = list(x.shape[:lead_nd]), x.shape[lead_nd:]
    # OCaml's `/` truncates toward zero (a kernel wider than the padded extent gives 0 or 1 windows)
    out_spatial = [int(((spatial[i] + padding[i][0] + padding[i][1]) - (dilation[i] * (kernel_size[i] - 1) + 1))
                       / stride[i]) + 1 for i in range(k)]
    kp, l = 1, 1
    for v in kernel_size:
        kp *= v
    for v in out_spatial:
        l *= v
    out = _create(x.context, x.dtype, leading + [kp, l])
    flat = [v for pr in padding for v in pr]
    do, dx = out._desc(), x._desc()
    _call(x.context, "unfold", x.context._lib.nxc_unfold, ctypes.byref(do), ctypes.byref(dx), k, _i64(kernel_size),
          _i64(stride), _i64(dilation), _i64(flat))
    return out


def fold(x: Tensor, output_size, kernel_size, stride, dilation, padding) -> Tensor:
    k = len(kernel_size)
    leading = list(x.shape[:len(x.shape) - 2])
    out = _create(x.context, x.dtype, leading + list(output_size))
    flat = [v for pr in padding for v in pr]
    do, dx = out._desc(), x._desc()
    _call(x.context, "fold", x.context._lib.nxc_fold, ctypes.byref(do), ctypes.byref(dx), k, _i64(output_size),
          _i64(kernel_size), _i64(stride), _i64(dilation), _i64(flat))
    return out


# ---- sort family (backend_c/nx_backend.ml:321-338) ---------------------------------------
def sort(x: Tensor, axis: int, descending: bool = False) -> Tensor:
    out = _create(x.context, x.dtype, x.shape)
    do, dx = out._desc(), x._desc()
    _call(x.context, "sort", x.context._lib.nxc_sort, 0, ctypes.byref(do), ctypes.byref(dx), int(axis),
          1 if descending else 0)
    return out


def argsort(x: Tensor, axis: int, descending: bool = False) -> Tensor:
    out = _create(x.context, _dt.int32, x.shape)
    do, dx = out._desc(), x._desc()
    _call(x.context, "argsort", x.context._lib.nxc_sort, 1, ctypes.byref(do), ctypes.byref(dx), int(axis),
          1 if descending else 0)
    return out


# ---- matmul (backend_c/nx_backend.ml:485-500) ---------------------------------------------
def matmul(x: Tensor, y: Tensor) -> Tensor:
    xs, ys = x.shape, y.shape
    xnd, ynd = len(xs), len(ys)
    m, n = xs[xnd - 2], ys[ynd - 1]
    max_nd = xnd if xnd > ynd else ynd
    batch = []
    for i in range(max_nd - 2):
        ai, bi = i - (max_nd - xnd), i - (max_nd - ynd)
        sa = xs[ai] if ai >= 0 else 1
        sb = ys[bi] if bi >= 0 else 1
        batch.append(sa if sa > sb else sb)
    out = _create(x.context, x.dtype, batch + [m, n])
    do, dx, dy = out._desc(), x._desc(), y._desc()
    _call(x.context, "matmul", x.context._lib.nxc_matmul, ctypes.byref(do), ctypes.byref(dx),
          ctypes.byref(dy))
    return out


# ---- fft family (backend_c/nx_backend.ml:502-549) -----------------------------------------
# Unnormalised transforms; the binding owns the output shape and dtype, the engine reads only
# the last entry of `s`.
def _axes_arg(axes):
    axes = [int(a) for a in axes]
    return (ctypes.c_int * _b.max(len(axes), 1))(*axes), len(axes)


def _fft(x: Tensor, axes, inverse: bool, opname: str) -> Tensor:
    out = _create(x.context, x.dtype, x.shape)
    ax, n = _axes_arg(axes)
    if n == 0 and out.buffer.nbytes:
        # nothing is transformed and nothing written: the reference's calloc'ed output stays zero
        check(x.context.ptr, opname, x.context._lib.nxc_memset(x.context.ptr, out.buffer.ptr, 0, out.buffer.nbytes))
    do, dx = out._desc(), x._desc()
    _call(x.context, opname, x.context._lib.nxc_fft, ctypes.byref(do), ctypes.byref(dx), ax, n, 1 if inverse else 0)
    return out


def fft(x: Tensor, axes) -> Tensor:
    return _fft(x, axes, False, "fft")


def ifft(x: Tensor, axes) -> Tensor:
    return _fft(x, axes, True, "ifft")


def rfft(x: Tensor, dtype, axes) -> Tensor:
    axes = [int(a) for a in axes]
    last = axes[-1]
    shape = list(x.shape)
    shape[last] = x.shape[last] // 2 + 1
    out = _create(x.context, _dt.of(dtype), shape)
    ax, n = _axes_arg(axes)
    do, dx = out._desc(), x._desc()
    _call(x.context, "rfft", x.context._lib.nxc_rfft, ctypes.byref(do), ctypes.byref(dx), ax, n)
    return out


def irfft(x: Tensor, dtype, axes, s=None) -> Tensor:
    axes = [int(a) for a in axes]
    last = axes[-1]
    size = int(s[-1]) if s is not None else (x.shape[last] - 1) * 2
    shape = list(x.shape)
    shape[last] = size
    out = _create(x.context, _dt.of(dtype), shape)
    ax, n = _axes_arg(axes)
    do, dx = out._desc(), x._desc()
    _call(x.context, "irfft", x.context._lib.nxc_irfft, ctypes.byref(do), ctypes.byref(dx), ax, n,
          int(s[-1]) if s is not None else 0)
    return out


# ---- linalg tier 1 (backend_c/nx_backend.ml:551-625) ---------------------------------------
class LinalgError(Failure):
    """`Backend_intf.Linalg_error {op; kind}` (backend_intf.ml:6-33): a numeric failure, lifted
    from the engine's Failure text exactly as the reference veneer's `reraise_linalg` does."""

    def __init__(self, op, kind, text):
        super().__init__(text)
        self.op, self.kind = op, kind


def _reraise_linalg(op, fn):
    try:
        return fn()
    except Failure as e:
        msg = str(e)
        for suffix, kind in (("matrix is not positive definite", "Not_positive_definite"),
                             ("triangular matrix is singular", "Singular"),
                             ("eigenvalue iteration did not converge", "No_convergence")):
            if msg.endswith(suffix):
                raise LinalgError(op, kind, msg) from None
        raise


def cholesky(x: Tensor, upper: bool = False) -> Tensor:
    out = _create(x.context, x.dtype, x.shape)
    do, dx = out._desc(), x._desc()
    _reraise_linalg("cholesky", lambda: _call(x.context, "cholesky", x.context._lib.nxc_cholesky,
                                              ctypes.byref(do), ctypes.byref(dx), 1 if upper else 0))
    return out


def triangular_solve(a: Tensor, b: Tensor, upper: bool = False, transpose: bool = False,
                     unit_diag: bool = False) -> Tensor:
    vector_rhs = len(b.shape) == len(a.shape) - 1
    bm = reshape(b, tuple(b.shape) + (1,)) if vector_rhs else b
    out = _create(b.context, b.dtype, bm.shape)
    flags = (1 if upper else 0) | (2 if transpose else 0) | (4 if unit_diag else 0)
    do, da, db = out._desc(), a._desc(), bm._desc()
    _reraise_linalg("triangular_solve", lambda: _call(b.context, "triangular_solve", b.context._lib.nxc_triangular_solve,
                                                      ctypes.byref(do), ctypes.byref(da), ctypes.byref(db), flags))
    return reshape(out, b.shape) if vector_rhs else out


def qr(x: Tensor, reduced: bool = True):
    s = list(x.shape)
    m, n = s[-2], s[-1]
    k = m if m < n else n
    qs, rs = list(s), list(s)
    if reduced:
        qs[-1], rs[-2] = k, k
    else:
        qs[-1] = m
    q, r = _create(x.context, x.dtype, qs), _create(x.context, x.dtype, rs)
    dq, dr, dx = q._desc(), r._desc(), x._desc()
    _reraise_linalg("qr", lambda: _call(x.context, "qr", x.context._lib.nxc_qr, ctypes.byref(dq), ctypes.byref(dr),
                                        ctypes.byref(dx), 1 if reduced else 0))
    return q, r


# ---- linalg tier 2: eigh / eigvalsh (backend_c/nx_backend.ml:627-648) ------------------------
def _eigh_values(x: Tensor) -> Tensor:
    return _create(x.context, _dt.float64, tuple(x.shape[:-2]) + (x.shape[-1],))


def eigvalsh(x: Tensor) -> Tensor:
    w = _eigh_values(x)
    dw, dx = w._desc(), x._desc()
    _reraise_linalg("eigvalsh", lambda: _call(x.context, "eigvalsh", x.context._lib.nxc_eigh, ctypes.byref(dw),
                                              ctypes.byref(dx), ctypes.byref(dx), 0))
    return w


def eigh(x: Tensor):
    w = _eigh_values(x)
    v = _create(x.context, x.dtype, x.shape)
    dw, dv, dx = w._desc(), v._desc(), x._desc()
    _reraise_linalg("eigh", lambda: _call(x.context, "eigh", x.context._lib.nxc_eigh, ctypes.byref(dw),
                                          ctypes.byref(dv), ctypes.byref(dx), 1))
    return w, v


# ---- linalg tier 3: svd, eig / eigvals (backend_c/nx_backend.ml:650-707) ----------------------
def svd(x: Tensor, full_matrices: bool = False):
    """(U, S, V^H); S is float64 [batch..., k]; thin gives U m x k and V^H k x n, full gives m x m and
    n x n -- the flag travels in the output shapes (backend_c/nx_backend.ml:653-677)."""
    sh = list(x.shape)
    m, n = sh[-2], sh[-1]
    k = m if m < n else n
    batch = sh[:-2]
    u = _create(x.context, x.dtype, batch + ([m, m] if full_matrices else [m, k]))
    s = _create(x.context, _dt.float64, batch + [k])
    vt = _create(x.context, x.dtype, batch + ([n, n] if full_matrices else [k, n]))
    du, ds, dv, dx = u._desc(), s._desc(), vt._desc(), x._desc()
    _reraise_linalg("svd", lambda: _call(x.context, "svd", x.context._lib.nxc_svd, ctypes.byref(du), ctypes.byref(ds),
                                         ctypes.byref(dv), ctypes.byref(dx)))
    return u, s, vt


def _eig_values(x: Tensor) -> Tensor:
    return _create(x.context, _dt.complex128, tuple(x.shape[:-2]) + (x.shape[-1],))


def eigvals(x: Tensor) -> Tensor:
    """Eigenvalues only, always complex128; the stub raises as "eig" for both entry points
    (nx_c_eig.c:1310-1326) and the values-only call passes `w` in the eigenvector slot."""
    w = _eig_values(x)
    dw, dx = w._desc(), x._desc()
    _reraise_linalg("eigvals", lambda: _call(x.context, "eig", x.context._lib.nxc_eig, ctypes.byref(dw),
                                             ctypes.byref(dw), ctypes.byref(dx), 0))
    return w


def eig(x: Tensor):
    w = _eig_values(x)
    v = _create(x.context, _dt.complex128, x.shape)
    dw, dv, dx = w._desc(), v._desc(), x._desc()
    _reraise_linalg("eig", lambda: _call(x.context, "eig", x.context._lib.nxc_eig, ctypes.byref(dw), ctypes.byref(dv),
                                         ctypes.byref(dx), 1))
    return w, v
