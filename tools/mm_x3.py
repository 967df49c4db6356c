#!/usr/bin/env python
"""f32 matmul accuracy and speed on one B200: the exact CUDA-core kernel ("f32"), the 3xTF32
tensor-core path ("f32x3") and plain tf32, against a float64 product of the same f32 inputs.
Error is reported relative to max|C| and to max (|A||B|) (the quantity a dot product's rounding
error scales with); inputs: N(0,1) and uniform [0,1) (all positive: the worst case for a biased
accumulator). Prints one JSON object (profiles/mm_x3_rNN.json)."""
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


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


rows = []
rng = np.random.default_rng(0)
for (m, k, n) in ((1024, 1024, 1024), (512, 8192, 512), (2048, 2048, 2048)):
    for dist in ("normal", "positive"):
        a = (rng.standard_normal((m, k)) if dist == "normal" else rng.random((m, k))).astype(np.float32)
        b = (rng.standard_normal((k, n)) if dist == "normal" else rng.random((k, n))).astype(np.float32)
        want = a.astype(np.float64) @ b.astype(np.float64)
        bound = float((np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)).max())
        ta, tb = B.from_host(ctx, a.reshape(-1)), B.from_host(ctx, b.reshape(-1))
        ta, tb = B.reshape(ta, [m, k]), B.reshape(tb, [k, n])
        bt = B.permute(B.contiguous(B.permute(tb, [1, 0])), [1, 0])  # same values, transposed storage
        r = {"m": m, "k": k, "n": n, "inputs": dist}
        for mode in ("f32", "f32x3", "tf32"):
            ctx.set_matmul_mode(mode)
            for lay, y in (("nn", tb), ("nt", bt)):
                got = B.to_numpy(B.matmul(ta, y)).astype(np.float64)
                err = float(np.abs(got - want).max())
                r[f"{mode}_{lay}_err_rel_maxC"] = err / float(np.abs(want).max())
                r[f"{mode}_{lay}_err_rel_absprod"] = err / bound
        rows.append(r)
        print(r, file=sys.stderr)
speed = []
for M in (2048, 4096, 8192):
    x = B.reshape(B.from_host(ctx, rng.standard_normal(M * M).astype(np.float32)), [M, M])
    y = B.reshape(B.from_host(ctx, rng.standard_normal(M * M).astype(np.float32)), [M, M])
    r = {"M": M}
    for mode in ("f32", "f32x3", "tf32"):
        ctx.set_matmul_mode(mode)
        for lay, (p, q) in (("nn", (x, y)), ("tn", (B.permute(x, [1, 0]), y))):
            ms = timeit(lambda: B.matmul(p, q), reps=3 if (mode == "f32" and M >= 8192) else 5)
            r[f"{mode}_{lay}_ms"] = round(ms, 3)
            r[f"{mode}_{lay}_tflops"] = round(2.0 * M ** 3 / ms / 1e9, 1)
    speed.append(r)
    print(r, file=sys.stderr)
ctx.set_matmul_mode("f32")
print(json.dumps({"accuracy": rows, "speed": speed}))
