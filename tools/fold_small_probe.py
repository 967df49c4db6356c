#!/usr/bin/env python
"""Device time of the leading-axis sums a data-parallel / batched-gradient step issues (the
unbroadcast after a batched matmul: [B, C, N] -> [C, N]) and a few neighbours, replayed from a
captured graph so that no host time is in the number. One JSON object."""
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
rng = np.random.default_rng(0)
rows = []
CASES = [((4, 768, 3072), [0]), ((4, 768, 768), [0]), ((4, 3072, 768), [0]), ((8, 768, 3072), [0]), ((256, 3072), [0]),
         ((256, 768), [0]), ((4, 12, 64, 64), [3]), ((256, 50257), [1]), ((64, 768, 3072), [0])]
for shape, axes in CASES:
    n = int(np.prod(shape))
    x = B.reshape(B.from_host(ctx, rng.standard_normal(n).astype(np.float32)), list(shape))
    for _ in range(3):
        y = B.reduce(x, "sum", axes)
    with ctx.capture() as g:
        for _ in range(20):
            y = B.reduce(x, "sum", axes)
    g.launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        g.launch()
    e1.record(stream)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 100 * 1e3
    kernels = g.kernels // 20
    g.close()
    out_n = n // int(np.prod([shape[a] for a in axes]))
    byts = 4 * (n + out_n)
    rows.append({"shape": list(shape), "axes": axes, "us": round(us, 2), "kernels": kernels, "gbs": round(byts / us / 1e3, 1)})
    print(rows[-1], file=sys.stderr)
print(json.dumps({"rows": rows}))
