#!/usr/bin/env python
"""bf16 NN GEMM at 4096^3 / 8192^3 under the tuning knobs of the tensor-core kernel (tile width,
rasterisation group, single-CTA vs pair): which resource binds? One JSON object."""
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
out = {}
for n in (2048, 4096, 8192, 16384):
    rng = np.random.default_rng(n)
    a = B.cast(B.reshape(B.from_host(ctx, rng.standard_normal(n * n).astype(np.float32)), [n, n]), "bf16")
    b = B.cast(B.reshape(B.from_host(ctx, rng.standard_normal(n * n).astype(np.float32)), [n, n]), "bf16")
    for name, env in (("default", {}), ("tile_n=128", {"NX_CUDA_MM_TILE_N": "128"}), ("tile_n=64", {"NX_CUDA_MM_TILE_N": "64"}),
                      ("group=4", {"NX_CUDA_MM_GROUP": "4"}), ("group=16", {"NX_CUDA_MM_GROUP": "16"}),
                      ("group=32", {"NX_CUDA_MM_GROUP": "32"}), ("single-cta", {"NX_CUDA_MM_PAIR": "0"}),
                      ("wide 256x512", {"NX_CUDA_MM_WIDE": "1"}),
                      ("prefetch=7", {"NX_CUDA_MM_PREFETCH": "7"}), ("prefetch=10", {"NX_CUDA_MM_PREFETCH": "10"}),
                      ("prefetch=16", {"NX_CUDA_MM_PREFETCH": "16"}), ("prefetch=32", {"NX_CUDA_MM_PREFETCH": "32"})):
        for k in ("NX_CUDA_MM_TILE_N", "NX_CUDA_MM_GROUP", "NX_CUDA_MM_PAIR", "NX_CUDA_MM_WIDE", "NX_CUDA_MM_PREFETCH"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for _ in range(3):
            B.matmul(a, b)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(4):
                B.matmul(a, b)
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 4)
        ms = sorted(ts)[2]
        out[f"{n}/{name}"] = {"ms": round(ms, 4), "tflops": round(2.0 * n ** 3 / ms / 1e9, 1)}
print(json.dumps(out))
