#!/usr/bin/env python
"""Tensor-core GEMM tile-mode comparison on one B200: the single-CTA 128x256 kernel
(NX_CUDA_MM_PAIR=0) against the 2-CTA 256x256 kernel (=1) and the engine's own choice
(unset), bf16 / tf32, square and skinny shapes, NN / NT / TN. CUDA events, median of 10.
Prints one JSON object (profiles/mm_modes_rNN.json)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raven_b200.backend as B  # noqa: E402
from raven_b200 import dtype as D  # noqa: E402

pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
BF16 = pk["bf16_tflops"]
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = B.create_context(device=0, stream=stream.cuda_stream)


def timeit(fn, reps=10, warm=3):
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


def rand(shape, dt):
    n = int(np.prod(shape))
    blk = min(n, 1 << 22)
    h = np.random.default_rng(0).uniform(-1, 1, blk).astype(np.float32)
    t = B.from_host(ctx, h)
    if blk < n:
        t = B.reshape(B.contiguous(B.expand(B.reshape(t, [1, blk]), [n // blk, blk])), [n])
    t = B.reshape(t, shape)
    return t if dt is D.float32 else B.cast(t, dt)


rows = []
shapes = [(8192, 8192, 8192), (4096, 4096, 4096), (2048, 2048, 2048), (16384, 16384, 16384), (8192, 4096, 1024),
          (8192, 4096, 4096), (1024, 8192, 8192), (256, 8192, 8192)]
for dtn, mm in (("bf16", "f32"), ("f32", "tf32")):
    dt = D.of(dtn)
    ctx.set_matmul_mode(mm)
    for (m, k, n) in shapes:
        if dtn == "f32" and m * k * n > 8192 ** 3:
            continue
        a, b = rand([m, k], dt), rand([k, n], dt)
        at = B.permute(rand([k, m], dt), [1, 0])
        bt = B.permute(rand([n, k], dt), [1, 0])
        lays = {"NN": (a, b)} if (m, k, n) != (8192, 8192, 8192) else {"NN": (a, b), "NT": (a, bt), "TN": (at, b)}
        for lay, (x, y) in lays.items():
            r = {"dtype": dtn if mm == "f32" else "tf32", "m": m, "k": k, "n": n, "layout": lay}
            for mode in ("0", "1", None):
                if mode is None:
                    os.environ.pop("NX_CUDA_MM_PAIR", None)
                else:
                    os.environ["NX_CUDA_MM_PAIR"] = mode
                ms = timeit(lambda: B.matmul(x, y), reps=5 if m * k * n > 8192 ** 3 else 10)
                tf = 2.0 * m * k * n / (ms * 1e-3) / 1e12
                key = {"0": "single", "1": "pair", None: "auto"}[mode]
                r[key + "_ms"] = round(ms, 4)
                r[key + "_tflops"] = round(tf, 1)
            r["pair_frac_of_peak"] = round(r["pair_tflops"] / BF16, 3)
            rows.append(r)
        del a, b, at, bt
print(json.dumps({"peak_bf16_tflops": BF16, "rows": rows}))
