#!/usr/bin/env python
"""BASELINE.json configs[3]: Rune.grad of a 4-layer 4096-wide MLP loss on a synthetic batch of
8192, replayed as the eager primitive sequence Rune's reverse-mode handler emits
(packages/rune/lib/reverse.ml): forward h = relu(h W + b) per layer (matmul, add with a
broadcast bias, max with a scalar 0), MSE loss (sub, mul, sum); backward per layer
relu' = cast(cmplt 0 pre) * g (reverse.ml:164-171), dW = matmul(transpose h, g) and
dh = matmul(g, transpose W) with TRANSPOSED VIEWS straight into the GEMM (reverse.ml:585-654),
db = sum g over the batch axis (broadcast undo, reverse.ml:32-52). One backend call per op,
fresh output per op, nothing fused -- exactly what the eager path does.

Prints a JSON line with ms/step and GEMM TFLOP/s (2*B*W*W*3 per layer) for bf16 and f32(tf32)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raven_b200.backend as B  # noqa: E402
from raven_b200 import dtype as D  # noqa: E402

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = B.create_context(device=0, stream=stream.cuda_stream)
BATCH, WIDTH, LAYERS = int(os.environ.get("MLP_BATCH", 8192)), int(os.environ.get("MLP_WIDTH", 4096)), 4


def rand(shape, dt, scale):
    rng = np.random.default_rng(0)
    n = int(np.prod(shape))
    blk = min(n, 1 << 22)
    t = B.from_host(ctx, (rng.standard_normal(blk) * scale).astype(np.float32))
    if blk < n:
        t = B.reshape(B.contiguous(B.expand(B.reshape(t, [1, blk]), [n // blk, blk])), [n])
    t = B.reshape(t, shape)
    return t if dt is D.float32 else B.cast(t, dt)


def run(dt):
    Ws = [rand([WIDTH, WIDTH], dt, 1.0 / np.sqrt(WIDTH)) for _ in range(LAYERS)]
    bs = [rand([WIDTH], dt, 0.01) for _ in range(LAYERS)]
    x = rand([BATCH, WIDTH], dt, 1.0)
    y = rand([BATCH, WIDTH], dt, 1.0)
    zero = B.full(ctx, dt, [], 0.0)
    inv = B.full(ctx, dt, [], 1.0 / (BATCH * WIDTH))

    def step():
        hs, pres = [x], []
        h = x
        for W, b in zip(Ws, bs):                       # forward
            pre = B.add(B.matmul(h, W), B.expand(B.reshape(b, [1, WIDTH]), [BATCH, WIDTH]))
            h = B.max(pre, B.expand(zero, [BATCH, WIDTH]))
            pres.append(pre)
            hs.append(h)
        diff = B.sub(h, y)
        loss = B.mul(B.reduce(B.mul(diff, diff), "sum", [0, 1]), inv)
        g = B.mul(diff, B.expand(B.mul(inv, B.full(ctx, dt, [], 2.0)), [BATCH, WIDTH]))   # dloss/dh
        grads = []
        for li in range(LAYERS - 1, -1, -1):           # backward
            mask = B.cast(B.cmplt(B.expand(zero, [BATCH, WIDTH]), pres[li]), dt)
            g = B.mul(g, mask)
            dW = B.matmul(B.permute(hs[li], [1, 0]), g)
            db = B.reduce(g, "sum", [0])
            grads.append((dW, db))
            if li > 0:
                g = B.matmul(g, B.permute(Ws[li], [1, 0]))
        return loss, grads

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ctx.launch_count()
    e0.record(stream)
    reps = 5
    for _ in range(reps):
        loss, grads = step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * BATCH * WIDTH * WIDTH * (3 * LAYERS - 1)
    return {"ms_per_step": round(ms, 3), "gemm_tflops": round(flops / (ms * 1e-3) / 1e12, 1),
            "launches_per_step": (ctx.launch_count() - n0) // reps, "loss": float(np.asarray(B.to_numpy(B.cast(loss, D.float32))))}


out = {"workload": f"Rune.grad-shaped MLP replay: {LAYERS} layers x {WIDTH}, batch {BATCH}, relu, MSE, eager unfused"}
out["bf16"] = run(D.bfloat16)
ctx.set_matmul_mode("tf32")
out["f32_tf32"] = run(D.float32)
ctx.set_matmul_mode("f32")   # the default: f32-class accuracy, large products as 3xTF32 on the tensor cores
out["f32_default_3xtf32"] = run(D.float32)
if os.environ.get("MLP_EXACT", "0") == "1":
    ctx.set_matmul_mode("ieee")
    out["f32_ieee_cuda_cores"] = run(D.float32)
    ctx.set_matmul_mode("f32")
print(json.dumps(out))
