"""An oracle-backed stand-in for `raven_b200.backend` (tests only).

Same function names and argument meaning, but tensors are HostViews over numpy storage and every
compute op is answered by the oracle (the reference's own C backend when oracle/_ref is built,
else the C restatement). Movement ops are the same pure view rewrites (raven_b200.view). With it
a whole op SEQUENCE -- the tape of tools/tape.py, the GPT-2 step of tools/gpt2_step.py, the
sharded host logic -- runs on the CPU reference, so the CUDA replay can be held to the same
sequence computed by the reference, and the host logic can be tested without a GPU."""
from __future__ import annotations

import numpy as np

from oracle.hostview import HostView
from raven_b200 import dtype as D
from raven_b200.view import View, c_contiguous_strides, numel


class OT:
    """Tensor double: a HostView plus the attributes the callers read."""

    __slots__ = ("hv", "dtype", "shape", "strides", "offset", "context")

    def __init__(self, hv, context=None):
        self.hv, self.dtype, self.shape = hv, D.of(hv.dtype), tuple(hv.shape)
        self.strides, self.offset, self.context = tuple(hv.strides), hv.offset, context


class Ctx:
    """Context double."""

    def sync(self):
        pass

    def launch_count(self):
        return 0


class OracleBackend:
    def __init__(self, oracle=None):
        if oracle is None:
            from tests import harness as H
            oracle = H.get_oracle()
        self.o = oracle
        for op in ("neg recip abs sign sqrt exp log sin cos tan asin acos atan sinh cosh tanh trunc ceil floor "
                   "round erf").split():
            setattr(self, op, self._unary(op))
        for op in "add sub mul idiv fdiv max min pow atan2 xor".split():
            setattr(self, op, self._binary(op))
        for op in "cmpeq cmpne cmplt cmple".split():
            setattr(self, op, self._compare(op))

    def _unary(self, op):
        return lambda x: OT(self.o.unary(op, x.hv), x.context)

    def _binary(self, op):
        return lambda a, b: OT(self.o.binary(op, a.hv, b.hv), a.context)

    def _compare(self, op):
        return lambda a, b: OT(self.o.compare(op, a.hv, b.hv), a.context)

    # ---- creation / transfer -------------------------------------------------------------------
    def create_context(self, *a, **k):
        return Ctx()

    def full(self, ctx, dt, shape, value):
        dt = D.of(dt)
        n = numel(shape)
        return OT(HostView(np.repeat(D.encode_scalar(dt, value), n), dt.name, list(shape)), ctx)

    def buffer(self, ctx, dt, shape):
        return OT(HostView.empty(D.of(dt).name, list(shape)), ctx)

    def from_host(self, ctx, array, dt=None):
        a = np.ascontiguousarray(array).reshape(-1)
        if dt is None:
            name = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.int32): "i32",
                    np.dtype(np.int64): "i64", np.dtype(np.uint8): "u8"}[a.dtype]
        else:
            name = D.of(dt).name
        return OT(HostView(a.copy(), name, [a.size]), ctx)

    def to_numpy(self, t):
        return t.hv.numpy()

    def to_host(self, t):
        return t.hv.storage

    # ---- movement: the same view arithmetic as the product --------------------------------------
    def _view(self, t, v):
        return OT(HostView(t.hv.storage, t.hv.dtype, v.shape, v.strides, v.offset), t.context)

    def reshape(self, t, shape):
        return self._view(t, View(t.shape, t.strides, t.offset).reshape(shape))

    def expand(self, t, shape):
        return self._view(t, View(t.shape, t.strides, t.offset).expand(shape))

    def permute(self, t, axes):
        return self._view(t, View(t.shape, t.strides, t.offset).permute(axes))

    def shrink(self, t, bounds):
        return self._view(t, View(t.shape, t.strides, t.offset).shrink(bounds))

    def is_c_contiguous(self, t):
        return t.offset == 0 and tuple(t.strides) == tuple(c_contiguous_strides(t.shape))

    # ---- compute ----------------------------------------------------------------------------------
    def where(self, c, a, b):
        return OT(self.o.where(c.hv, a.hv, b.hv), a.context)

    def cast(self, x, dtype):
        return OT(self.o.cast(x.hv, D.of(dtype).name), x.context)

    def copy(self, x):
        return OT(self.o.copy(x.hv), x.context)

    def contiguous(self, x):
        return x if self.is_c_contiguous(x) else self.copy(x)

    def assign(self, dst, src):
        self.o.assign(dst.hv, src.hv)

    def reduce(self, x, op, axes):
        return OT(self.o.reduce(op, x.hv, sorted(axes)), x.context)

    def argmax(self, x, axis, keepdims=False):
        return OT(self.o.argreduce("argmax", x.hv, axis, keepdims), x.context)

    def argmin(self, x, axis, keepdims=False):
        return OT(self.o.argreduce("argmin", x.hv, axis, keepdims), x.context)

    def gather(self, data, idx, axis, trusted=False):
        return OT(self.o.gather(data.hv, idx.hv, axis), data.context)

    def scatter(self, template, idx, updates, axis, mode="set", unique_indices=False):
        return OT(self.o.scatter(template.hv, idx.hv, updates.hv, axis, mode), template.context)

    def cat(self, ts, axis):
        return OT(self.o.cat([t.hv for t in ts], axis), ts[0].context)

    def matmul(self, a, b):
        return OT(self.o.matmul(a.hv, b.hv), a.context)
