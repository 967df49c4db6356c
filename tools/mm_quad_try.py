#!/usr/bin/env python
"""The 4-CTA-cluster multicast GEMM (NX_CUDA_MM_QUAD=1) against the pair kernel: bit-identity of
the product (same accumulation order) over layouts and ragged shapes, then timing. One JSON object."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raven_b200.backend as B  # noqa: E402

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = B.create_context(device=0, stream=stream.cuda_stream)
out = {"identical": {}, "timing": {}}


def mk(shape, dt, seed):
    n = int(np.prod(shape))
    t = B.reshape(B.from_host(ctx, np.random.default_rng(seed).standard_normal(n).astype(np.float32)), list(shape))
    return B.cast(t, dt) if dt != "f32" else t


def run(a, b, quad):
    os.environ.pop("NX_CUDA_MM_QUAD", None)
    if quad:
        os.environ["NX_CUDA_MM_QUAD"] = "1"
    c = B.matmul(a, b)
    os.environ.pop("NX_CUDA_MM_QUAD", None)
    return c


cases = [("bf16", 1024, 256, 512), ("bf16", 768, 192, 300), ("bf16", 1280, 1024, 2048), ("bf16", 4096, 4096, 4096),
         ("f16", 1536, 512, 1024), ("bf16", 520, 64, 256)]
for dt, m, k, n in cases:
    a, b = mk([m, k], dt, 1), mk([k, n], dt, 2)
    at, bt = B.permute(mk([k, m], dt, 3), [1, 0]), B.permute(mk([n, k], dt, 4), [1, 0])
    for name, (x, y) in {"nn": (a, b), "tn": (at, b), "nt": (a, bt), "tt": (at, bt)}.items():
        want = B.to_numpy(run(x, y, False))
        got = B.to_numpy(run(x, y, True))
        out["identical"][f"{dt} {m}x{k}x{n} {name}"] = bool(np.array_equal(want, got))
if os.environ.get("MM_TF32", "1") == "1":
    ctx.set_matmul_mode("tf32")
    a, b = mk([1024, 512], "f32", 5), mk([512, 768], "f32", 6)
    out["identical"]["tf32 1024x512x768 nn"] = bool(np.array_equal(B.to_numpy(run(a, b, False)), B.to_numpy(run(a, b, True))))
    ctx.set_matmul_mode("f32")
    a, b = mk([2048, 1024], "f32", 7), mk([1024, 2048], "f32", 8)
    out["identical"]["f32 (3xTF32) 2048x1024x2048"] = bool(np.array_equal(B.to_numpy(run(a, b, False)), B.to_numpy(run(a, b, True))))
for n in (4096, 8192, 16384):
    a, b = mk([n, n], "bf16", n), mk([n, n], "bf16", n + 1)
    for quad in (False, True):
        for _ in range(3):
            run(a, b, quad)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(4):
                run(a, b, quad)
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 4)
        ms = sorted(ts)[2]
        out["timing"][f"{n} {'quad' if quad else 'pair'}"] = {"ms": round(ms, 4), "tflops": round(2.0 * n ** 3 / ms / 1e9, 1)}
    del a, b
print(json.dumps(out))
