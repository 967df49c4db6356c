#!/usr/bin/env python
"""K-skew experiment for the tcgen05 GEMM: NX_CUDA_MM_KSKEW in {1 (lockstep), 4, 8, 16, 32} at
4096^3 / 8192^3 / 16384^3 bf16, NN / NT / TN, CUDA events (median). Prints one JSON object."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raven_b200.backend as B  # noqa: E402
from raven_b200 import dtype as D  # noqa: E402

BF16 = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = B.create_context(device=0, stream=stream.cuda_stream)


def timeit(fn, reps=7, warm=2):
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
    return B.cast(B.reshape(t, shape), dt)


rows = []
skews = [int(x) for x in os.environ.get("KSKEWS", "1,4,8,16,32").split(",")]
groups = [int(x) for x in os.environ.get("MM_GROUPS", "").split(",") if x]
for M in (4096, 8192, 16384):
    x, y = rand([M, M], D.bfloat16), rand([M, M], D.bfloat16)
    lays = {"NN": (x, y), "NT": (x, B.permute(y, [1, 0])), "TN": (B.permute(x, [1, 0]), y)}
    for lay, (p, q) in lays.items():
        r = {"M": M, "layout": lay}
        for ks in skews:
            os.environ["NX_CUDA_MM_KSKEW"] = str(ks)
            ms = timeit(lambda: B.matmul(p, q))
            r[f"skew{ks}_tflops"] = round(2.0 * M ** 3 / (ms * 1e-3) / 1e12, 1)
        os.environ["NX_CUDA_MM_KSKEW"] = "1"
        for g in groups:   # M-blocks per rasterisation group (L2 reuse of the B panels)
            os.environ["NX_CUDA_MM_GROUP"] = str(g)
            ms = timeit(lambda: B.matmul(p, q))
            r[f"group{g}_tflops"] = round(2.0 * M ** 3 / (ms * 1e-3) / 1e12, 1)
        os.environ.pop("NX_CUDA_MM_GROUP", None)
        rows.append(r)
        print(r, file=sys.stderr)
    del x, y, lays
print(json.dumps({"peak_bf16_tflops": BF16, "rows": rows}))
