#!/usr/bin/env python
"""HBM rate of the 1- and 2-byte dtypes through the map kernels (the sweep covers f32 / f64 / i32):
contiguous add, transposed-view add, contiguous(transpose), row-broadcast add, cast to f32, sum over
the inner axis; 2^28 elements, CUDA events (median of 7). Prints one JSON object."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raven_b200.backend as B  # noqa: E402
from raven_b200 import dtype as D  # noqa: E402

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = B.create_context(device=0, stream=stream.cuda_stream)
n, r = 1 << 28, 1 << 14


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[3]


h = np.random.default_rng(0).uniform(-4, 4, 1 << 22).astype(np.float32)
base = B.from_host(ctx, h)
base = B.reshape(B.contiguous(B.expand(B.reshape(base, [1, 1 << 22]), [n >> 22, 1 << 22])), [n])
out = {}
for dt in (D.bfloat16, D.float16, D.int8, D.uint8):
    a = B.cast(base, dt)
    b = B.cast(base, dt)
    es = dt.itemsize
    A, Bm = B.reshape(a, [r, r]), B.reshape(b, [r, r])
    cases = {
        "add contiguous": (lambda: B.add(a, b), 3 * n * es),
        "add transposed rhs": (lambda: B.add(A, B.permute(Bm, [1, 0])), 3 * n * es),
        "contiguous(transpose)": (lambda: B.contiguous(B.permute(A, [1, 0])), 2 * n * es),
        "add row-broadcast": (lambda: B.add(A, B.expand(B.shrink(Bm, [(0, 1), (0, r)]), [r, r])), (2 * n + r) * es),
        "cast->f32": (lambda: B.cast(a, D.float32), n * (es + 4)),
        "sum inner": (lambda: B.reduce(A, "sum", [1]), n * es),
    }
    res = {}
    for name, (fn, nbytes) in cases.items():
        ms = timeit(fn)
        res[name] = {"ms": round(ms, 4), "gbs": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / HBM, 3)}
    out[dt.name] = res
    print(dt.name, {k: v["frac"] for k, v in res.items()}, file=sys.stderr)
    del a, b, A, Bm
print(json.dumps(out))
