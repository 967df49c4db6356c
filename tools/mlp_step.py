#!/usr/bin/env python
"""BASELINE.json configs[3]: `Rune.grad` of a 4-layer 4096-wide MLP loss on a synthetic batch of
8192, run through the backend as the eager primitive sequence Rune's reverse-mode handler emits
(tools/tape.py restates packages/rune/lib/reverse.ml): forward h = relu(h W + b) per layer
(matmul, add with a broadcast bias, max with a scalar 0), mean-squared-error loss; backward the
pull-backs in reverse -- relu through `cast (greater a b)` masks (reverse.ml:164-171), dW =
matmul(transpose h, g) and dh = matmul(g, transpose W) with TRANSPOSED VIEWS straight into the
GEMM (reverse.ml:585-654), db by un-broadcasting (reverse.ml:32-52). One backend call per op,
fresh output per op, nothing fused. Timed eager (issued op by op) and as a captured step.

    python tools/mlp_step.py        # one JSON line: ms/step and GEMM TFLOP/s for bf16 and f32"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from tools.tape import Tape  # noqa: E402

BATCH, WIDTH, LAYERS = int(os.environ.get("MLP_BATCH", 8192)), int(os.environ.get("MLP_WIDTH", 4096)), 4


def _rand(B, ctx, shape, dt, scale, seed):
    from raven_b200 import dtype as D
    rng = np.random.default_rng(seed)
    n = int(np.prod(shape))
    blk = min(n, 1 << 22)
    t = B.from_host(ctx, (rng.standard_normal(blk) * scale).astype(np.float32))
    if blk < n:
        t = B.reshape(B.contiguous(B.expand(B.reshape(t, [1, blk]), [n // blk, blk])), [n])
    t = B.reshape(t, shape)
    return t if dt is D.float32 else B.cast(t, dt)


def value_and_grad(B, ctx, Ws, bs, x, y):
    tp = Tape(B, ctx)
    lw, lb = [tp.leaf(w) for w in Ws], [tp.leaf(b) for b in bs]
    h = tp.const(x)
    for w, b in zip(lw, lb):
        h = tp.relu(tp.linear(h, w, b))
    diff = tp.sub(h, tp.const(y))
    loss = tp.mean(tp.mul(diff, diff), [0, 1])
    tp.backward(loss, B.full(ctx, loss.dtype, [], 1.0))
    return loss.t, [tp.grad_of(v) for v in lw], [tp.grad_of(v) for v in lb]


def run_dtype(B, ctx, stream, dt, reps=5):
    import torch
    Ws = [_rand(B, ctx, [WIDTH, WIDTH], dt, 1.0 / np.sqrt(WIDTH), 10 + i) for i in range(LAYERS)]
    bs = [_rand(B, ctx, [WIDTH], dt, 0.01, 20 + i) for i in range(LAYERS)]
    x, y = _rand(B, ctx, [BATCH, WIDTH], dt, 1.0, 1), _rand(B, ctx, [BATCH, WIDTH], dt, 1.0, 2)

    def step():
        return value_and_grad(B, ctx, Ws, bs, x, y)

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for _ in range(2):
        step()
    n0 = ctx.launch_count()
    loss, _, _ = step()
    per = ctx.launch_count() - n0
    eager_ms = timed(step)
    with ctx.capture() as g:
        outs = step()
    g.launch()
    graph_ms = timed(g.launch)
    lg = float(np.asarray(B.to_numpy(B.cast(outs[0], "f32"))))
    le = float(np.asarray(B.to_numpy(B.cast(loss, "f32"))))
    g.close()
    flops = 2.0 * BATCH * WIDTH * WIDTH * (3 * LAYERS - 1)
    return {"ms_per_step": round(graph_ms, 3), "eager_ms_per_step": round(eager_ms, 3),
            "gemm_tflops": round(flops / (graph_ms * 1e-3) / 1e12, 1), "launches_per_step": int(per),
            "loss": le, "captured_loss_equals_eager": lg == le}


def run(ctx=None, stream=None):
    import torch

    import raven_b200.backend as B
    from raven_b200 import dtype as D
    if ctx is None:
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        ctx = B.create_context(device=int(os.environ.get("LOCAL_RANK", "0")), stream=stream.cuda_stream)
    out = {"workload": f"Rune.grad-shaped MLP: {LAYERS} layers x {WIDTH}, batch {BATCH}, relu, MSE; value_and_grad, eager "
                       f"primitives, unfused; ms_per_step = captured step replayed"}
    out["bf16"] = run_dtype(B, ctx, stream, D.bfloat16)
    out["f32_default_3xtf32"] = run_dtype(B, ctx, stream, D.float32)
    return out


if __name__ == "__main__":
    print(json.dumps(run()))
