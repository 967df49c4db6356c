"""A reverse-mode tape over the backend primitives -- the op sequence `Rune.grad` emits.

Rune differentiates by intercepting the backend-level effects (one per `Backend_intf.S` call)
and recording a pull-back per op (reference: packages/rune/lib/reverse.ml:106-660); the backward
pass replays the records in reverse, accumulating cotangents with `add`. Nothing is fused and
nothing is compiled: what reaches the backend is a long sequence of eager primitives. This module
restates those rules over any backend MODULE with the `raven_b200.backend` surface (the CUDA
backend, or tests/backend_double.py over the CPU oracle), so that the configs BASELINE.json names
above single ops -- the Rune.grad MLP loss (configs[3]) and the Kaun GPT-2 training step
(configs[4]) -- can be run through the backend exactly as the reference's frontend would drive it.

The frontend-level helpers (`bcast`, `sum`, `mean`, `softmax`, `log_softmax`, `layer_norm`,
`gelu_approx`, `linear`, `attention`, `cross_entropy_sparse`) expand into primitives the way
Nx's frontend and Kaun's layers do; each cites the source it follows. Where the reference
recomputes a value only to read its SHAPE (reverse.ml:372-376), the shape is computed
arithmetically instead -- no primitive is dropped whose result is used.

Test / workload infrastructure: the product is the backend underneath."""
from __future__ import annotations

import math


class V:
    """A traced value: the backend tensor and whether a cotangent flows to it."""

    __slots__ = ("t", "tracked", "uid", "is_leaf")
    _next = 0

    def __init__(self, t, tracked=False, is_leaf=False):
        self.t = t
        self.tracked = tracked
        self.is_leaf = is_leaf
        V._next += 1
        self.uid = V._next

    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def dtype(self):
        return self.t.dtype


class Tape:
    def __init__(self, backend, ctx):
        self.B = backend
        self.ctx = ctx
        self.records = []      # closures, run in reverse
        self.cot = {}          # uid -> cotangent tensor
        self.uses = {}         # leaf uid -> pull-backs still to run that feed it
        # called as on_leaf_final(leaf, cotangent) the moment a leaf's cotangent is complete, i.e.
        # while the rest of the backward pass is still to be issued (gradient buckets start their
        # all-reduce under it)
        self.on_leaf_final = None

    # ---- bookkeeping (reverse.ml:68-100: pull1 / pull2 / accumulate) --------------------------
    def leaf(self, t):
        return V(t, True, True)

    def const(self, t):
        return V(t, False)

    def _accumulate(self, x: V, g):
        if x.uid in self.cot:
            self.cot[x.uid] = self.B.add(self.cot[x.uid], g)
        else:
            self.cot[x.uid] = g

    def _pull(self, out: V, inputs, fns):
        """Record: cotangent of inputs[i] += fns[i](cotangent of out), for the tracked inputs."""
        live = [(x, f) for x, f in zip(inputs, fns) if x.tracked]
        if not live:
            return out
        out.tracked = True
        for x, _ in live:
            if x.is_leaf:
                self.uses[x.uid] = self.uses.get(x.uid, 0) + 1

        def rec():
            g = self.cot.pop(out.uid, None)
            for x, f in live:
                if g is not None:
                    self._accumulate(x, f(g))
                if x.is_leaf:
                    self.uses[x.uid] -= 1
                    if self.uses[x.uid] == 0 and self.on_leaf_final is not None and x.uid in self.cot:
                        self.on_leaf_final(x, self.cot[x.uid])
        self.records.append(rec)
        return out

    def backward(self, out: V, seed):
        self.cot[out.uid] = seed
        for rec in reversed(self.records):
            rec()
        self.records = []

    def grad_of(self, x: V):
        return self.cot.get(x.uid)

    # ---- helpers that are not differentiated ----------------------------------------------------
    def scalar(self, dtype, value):
        return self.B.full(self.ctx, dtype, [], value)

    def scalar_like(self, x: V, value) -> V:
        """`scalar_like x v` broadcast to x's shape (frontend.ml: reshape to rank-many 1s + expand)."""
        s = self.scalar(x.dtype, value)
        nd = len(x.shape)
        return V(self.B.expand(self.B.reshape(s, [1] * nd), list(x.shape)) if nd else s)

    def _unbroadcast(self, g, src_shape):
        """reverse.ml:32-52: sum the cotangent over the axes broadcasting added or stretched."""
        B = self.B
        dst = tuple(g.shape)
        src_shape = tuple(src_shape)
        if src_shape == dst:
            return g
        lead = len(dst) - len(src_shape)
        axes = list(range(lead)) + [i + lead for i, s in enumerate(src_shape) if s == 1 and dst[i + lead] > 1]
        if not axes:
            return g
        summed = B.reduce(g, "sum", axes)
        return B.reshape(summed, list(src_shape))

    def _broadcast_kept(self, t, shape_in, axes):
        """`broadcast_to shape_in (reshape kept t)` for a tensor reduced over `axes`."""
        kept = [1 if i in axes else d for i, d in enumerate(shape_in)]
        return self.B.expand(self.B.reshape(t, kept), list(shape_in))

    # ---- primitives and their pull-backs -------------------------------------------------------
    def add(self, a: V, b: V) -> V:   # reverse.ml:145
        return self._pull(V(self.B.add(a.t, b.t)), (a, b), (lambda g: g, lambda g: g))

    def sub(self, a: V, b: V) -> V:   # reverse.ml:146
        return self._pull(V(self.B.sub(a.t, b.t)), (a, b), (lambda g: g, lambda g: self.B.neg(g)))

    def mul(self, a: V, b: V) -> V:   # reverse.ml:147-150
        B = self.B
        return self._pull(V(B.mul(a.t, b.t)), (a, b), (lambda g: B.mul(g, b.t), lambda g: B.mul(g, a.t)))

    def div(self, a: V, b: V) -> V:   # reverse.ml:151-156
        B = self.B
        return self._pull(V(B.fdiv(a.t, b.t)), (a, b),
                          (lambda g: B.fdiv(g, b.t), lambda g: B.mul(B.neg(g), B.fdiv(a.t, B.mul(b.t, b.t)))))

    def neg(self, a: V) -> V:         # reverse.ml:192
        return self._pull(V(self.B.neg(a.t)), (a,), (lambda g: self.B.neg(g),))

    def exp(self, a: V) -> V:         # reverse.ml:228-232
        out = V(self.B.exp(a.t))
        return self._pull(out, (a,), (lambda g: self.B.mul(g, out.t),))

    def log(self, a: V) -> V:         # reverse.ml:233-235
        return self._pull(V(self.B.log(a.t)), (a,), (lambda g: self.B.mul(g, self.B.recip(a.t)),))

    def sqrt(self, a: V) -> V:        # reverse.ml:236-240, derivs.ml:18-19
        B = self.B
        out = V(B.sqrt(a.t))

        def pb(g):
            one, two = self.scalar_like(out, 1.0).t, self.scalar_like(out, 2.0).t
            return B.mul(g, B.fdiv(one, B.mul(two, out.t)))
        return self._pull(out, (a,), (pb,))

    def tanh(self, a: V) -> V:        # reverse.ml:223-227, derivs.ml:45
        B = self.B
        out = V(B.tanh(a.t))
        return self._pull(out, (a,), (lambda g: B.mul(g, B.sub(self.scalar_like(out, 1.0).t, B.mul(out.t, out.t))),))

    def max(self, a: V, b: V) -> V:   # reverse.ml:164-171
        B = self.B
        out = V(B.max(a.t, b.t))

        def mask(g):
            return B.cast(B.cmplt(b.t, a.t), g.dtype)   # greater a b
        return self._pull(out, (a, b), (lambda g: B.mul(g, mask(g)),
                                        lambda g: B.mul(g, B.sub(self.scalar_like(V(g), 1.0).t, mask(g)))))

    def where(self, cond, a: V, b: V) -> V:   # reverse.ml:253-274
        B = self.B
        out = V(B.where(cond, a.t, b.t))

        def pa(g):
            return self._unbroadcast(B.mul(g, B.cast(cond, g.dtype)), a.shape)

        def pb(g):
            m = B.cast(cond, g.dtype)
            return self._unbroadcast(B.mul(g, B.sub(self.scalar_like(V(g), 1.0).t, m)), b.shape)
        return self._pull(out, (a, b), (pa, pb))

    def contiguous(self, a: V) -> V:          # reverse.ml:353-354
        return self._pull(V(self.B.contiguous(a.t)), (a,), (lambda g: g,))

    def reshape(self, a: V, shape) -> V:      # reverse.ml:276-280
        """The frontend's reshape: a view when the strides allow it, else `contiguous` first."""
        src = list(a.shape)
        try:
            out = self.B.reshape(a.t, list(shape))
        except ValueError:
            a = self.contiguous(a)
            out = self.B.reshape(a.t, list(shape))

        def pb(g):
            try:
                return self.B.reshape(g, src)
            except ValueError:
                return self.B.reshape(self.B.contiguous(g), src)
        return self._pull(V(out), (a,), (pb,))

    def permute(self, a: V, axes) -> V:       # reverse.ml:281-287
        inv = [0] * len(axes)
        for i, d in enumerate(axes):
            inv[d] = i
        return self._pull(V(self.B.permute(a.t, list(axes))), (a,), (lambda g: self.B.permute(g, inv),))

    def expand(self, a: V, shape) -> V:       # reverse.ml:288-292
        src = a.shape
        return self._pull(V(self.B.expand(a.t, list(shape))), (a,), (lambda g: self._unbroadcast(g, src),))

    def cast(self, a: V, dtype) -> V:         # reverse.ml:348-352
        src = a.dtype
        return self._pull(V(self.B.cast(a.t, dtype)), (a,), (lambda g: self.B.cast(g, src),))

    def reduce_sum(self, a: V, axes) -> V:    # reverse.ml:357-366
        axes = sorted(axes)
        shape_in = a.shape
        return self._pull(V(self.B.reduce(a.t, "sum", axes)), (a,),
                          (lambda g: self._broadcast_kept(g, shape_in, axes),))

    def reduce_max(self, a: V, axes) -> V:    # reverse.ml:367-382
        B = self.B
        axes = sorted(axes)
        shape_in = a.shape
        out = V(B.reduce(a.t, "max", axes))

        def pb(g):
            mask = B.cast(B.cmpeq(a.t, self._broadcast_kept(out.t, shape_in, axes)), out.dtype)
            return B.mul(self._broadcast_kept(g, shape_in, axes), mask)
        return self._pull(out, (a,), (pb,))

    def gather(self, data: V, indices, axis) -> V:   # reverse.ml:521-526
        B = self.B

        def pb(g):
            zeros = B.full(self.ctx, data.dtype, list(data.shape), 0.0)
            return B.scatter(zeros, indices, g, axis, mode="add")
        return self._pull(V(B.gather(data.t, indices, axis)), (data,), (pb,))

    def matmul(self, a: V, b: V) -> V:        # reverse.ml:585-657
        B = self.B
        out = V(B.matmul(a.t, b.t))
        a_nd, b_nd = len(a.shape), len(b.shape)

        def t2(x):
            nd = len(x.shape)
            ax = list(range(nd))
            ax[-1], ax[-2] = ax[-2], ax[-1]
            return B.permute(x, ax)

        def pa(g):
            if a_nd == 2 and b_nd >= 3:
                gb = B.matmul(g, t2(b.t))
                return B.reduce(gb, "sum", list(range(len(g.shape) - 2)))
            return B.matmul(g, t2(b.t))

        def pb(g):
            if b_nd == 2 and a_nd >= 3:
                ag = B.matmul(t2(a.t), g)
                return B.reduce(ag, "sum", list(range(len(g.shape) - 2)))
            if a_nd == 2 and b_nd >= 3:
                at = t2(a.t)
                tgt = list(g.shape[:-2]) + list(at.shape)
                return B.matmul(B.expand(B.reshape(at, [1] + list(at.shape)), tgt), g)
            return B.matmul(t2(a.t), g)
        return self._pull(out, (a, b), (pa, pb))

    # ---- frontend expansions -----------------------------------------------------------------------
    def bcast(self, a: V, shape) -> V:
        """`broadcast_to`: right-align with leading 1s, then expand (frontend.ml broadcast rules)."""
        shape = list(shape)
        if list(a.shape) == shape:
            return a
        nd = len(shape)
        if len(a.shape) < nd:
            a = self.reshape(a, [1] * (nd - len(a.shape)) + list(a.shape))
        return self.expand(a, shape)

    def sum(self, a: V, axes, keepdims=False) -> V:
        axes = sorted(ax % len(a.shape) for ax in axes)
        s = self.reduce_sum(a, axes)
        return self.reshape(s, [1 if i in axes else d for i, d in enumerate(a.shape)]) if keepdims else s

    def max_reduce(self, a: V, axes, keepdims=False) -> V:
        axes = sorted(ax % len(a.shape) for ax in axes)
        s = self.reduce_max(a, axes)
        return self.reshape(s, [1 if i in axes else d for i, d in enumerate(a.shape)]) if keepdims else s

    def mean(self, a: V, axes, keepdims=False) -> V:   # frontend.ml:698-707
        s = self.sum(a, axes, keepdims)
        n = 1
        for ax in axes:
            n *= a.shape[ax % len(a.shape)]
        return self.div(s, self.scalar_like(s, float(max(1, n))))

    def mul_s(self, a: V, value) -> V:
        return self.mul(a, self.scalar_like(a, value))

    def add_s(self, a: V, value) -> V:
        return self.add(a, self.scalar_like(a, value))

    def softmax(self, x: V) -> V:          # frontend.ml:4242-4252, last axis
        mx = self.max_reduce(x, [-1], keepdims=True)
        e = self.exp(self.sub(x, self.bcast(mx, x.shape)))
        return self.div(e, self.bcast(self.sum(e, [-1], keepdims=True), x.shape))

    def log_softmax(self, x: V) -> V:      # frontend.ml:4254-4266
        mx = self.max_reduce(x, [-1], keepdims=True)
        shifted = self.sub(x, self.bcast(mx, x.shape))
        log_den = self.log(self.sum(self.exp(shifted), [-1], keepdims=True))
        return self.sub(shifted, self.bcast(log_den, x.shape))

    def layer_norm(self, x: V, gamma: V, beta: V, eps=1e-5) -> V:   # kaun/lib/layer_norm.ml:43-68
        mu = self.mean(x, [-1], keepdims=True)
        xc = self.sub(x, self.bcast(mu, x.shape))
        var = self.mean(self.mul(xc, xc), [-1], keepdims=True)
        normalized = self.div(xc, self.bcast(self.sqrt(self.add_s(var, eps)), x.shape))
        return self.add(self.mul(normalized, self.bcast(gamma, x.shape)), self.bcast(beta, x.shape))

    def gelu_approx(self, x: V) -> V:      # kaun/lib/fn.ml:31-38
        x3 = self.mul(x, self.mul(x, x))
        inner = self.mul_s(self.add(x, self.mul_s(x3, 0.044715)), math.sqrt(2.0 / math.pi))
        return self.mul_s(self.mul(x, self.add_s(self.tanh(inner), 1.0)), 0.5)

    def linear(self, x: V, w: V, b: V = None) -> V:   # kaun/lib/linear.ml:50-52
        y = self.matmul(x, w)
        return y if b is None else self.add(y, self.bcast(b, y.shape))

    def relu(self, x: V) -> V:
        return self.max(x, self.scalar_like(x, 0.0))

    def attention(self, x: V, p, num_heads, mask) -> V:
        """kaun/lib/attention.ml: q/k/v projections, split heads ([B,T,C] -> [B,H,T,D] by reshape +
        transpose), scores = q k^T * scale, causal `where mask scores -inf`, softmax, probs v, merge
        heads, output projection. `mask` is a bool tensor broadcast to [B,H,T,T] (a view)."""
        Bt, T, C = x.shape
        D = C // num_heads

        def heads(t):
            return self.permute(self.reshape(t, [Bt, T, num_heads, D]), [0, 2, 1, 3])
        q, k, v = (heads(self.linear(x, p[n + ".w"], p[n + ".b"])) for n in ("q", "k", "v"))
        scores = self.mul_s(self.matmul(q, self.permute(k, [0, 1, 3, 2])), 1.0 / math.sqrt(D))
        scores = self.where(mask, scores, self.scalar_like(scores, float("-inf")))
        ctxv = self.matmul(self.softmax(scores), v)
        merged = self.reshape(self.permute(ctxv, [0, 2, 1, 3]), [Bt, T, C])
        return self.linear(merged, p["out.w"], p["out.b"])

    def cross_entropy_sparse(self, logits: V, labels) -> V:   # kaun/lib/loss.ml:64-78, `Mean
        lp = self.log_softmax(logits)
        idx = self.B.reshape(labels, list(labels.shape) + [1])
        picked = self.gather(lp, idx, len(lp.shape) - 1)
        nll = self.neg(self.reshape(picked, list(labels.shape)))
        return self.mean(nll, list(range(len(nll.shape))))
