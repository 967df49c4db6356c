#!/usr/bin/env python
"""The reduce / argreduce / strided-map layouts of the config-2 sweep that sit furthest from the
HBM roofline, one launch each after a warm-up, so `ncu -k regex:...` can capture them:
    tools/ncu_export.sh weak 'nxc_(fold|map)' 0 40 -- python tools/fold_cases.py
With FOLD_CASES_TIME=1 it prints CUDA-event timings instead (median of 10)."""
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
lg = int(os.environ.get("FOLD_CASES_LOG2", "28"))
n = 1 << lg
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def rand(dt, n):
    blk = min(n, 1 << 22)
    h = np.random.default_rng(0).uniform(-4, 4, blk).astype(dt.np)
    t = B.from_host(ctx, h)
    if blk < n:
        t = B.reshape(B.contiguous(B.expand(B.reshape(t, [1, blk]), [n // blk, blk])), [n])
    return t


DT = D.of(os.environ.get("FOLD_CASES_DT", "f32"))
ES = DT.itemsize
a = rand(DT, n)
b = rand(DT, n)
short = B.reshape(a, [n // 256, 256])
sq = B.reshape(a, [1 << (lg // 2), n >> (lg // 2)])
bs = B.reshape(b, [1 << (lg // 2), n >> (lg // 2)])
r, c = sq.shape
cases = {
    "sum inner [N/256,256]": (lambda: B.reduce(short, "sum", [1]), ES * n),
    "max inner [N/256,256]": (lambda: B.reduce(short, "max", [1]), ES * n),
    "argmax inner [N/256,256]": (lambda: B.argmax(short, 1), ES * n),
    "sum outer [N/256,256]": (lambda: B.reduce(short, "sum", [0]), ES * n),
    "argmax outer [N/256,256]": (lambda: B.argmax(short, 0), ES * n),
    "argmax outer sqrt": (lambda: B.argmax(sq, 0), ES * n),
    "argmax inner sqrt": (lambda: B.argmax(sq, 1), ES * n),
    "max outer sqrt": (lambda: B.reduce(sq, "max", [0]), ES * n),
    "add col-broadcast": (lambda: B.add(sq, B.expand(B.shrink(bs, [(0, r), (0, 1)]), [r, c])), 2 * ES * n),
    "add row-broadcast": (lambda: B.add(sq, B.expand(B.shrink(bs, [(0, 1), (0, c)]), [r, c])), 2 * ES * n),
    "add flipped": (lambda: B.add(sq, B.flip(bs, [True, True])), 3 * ES * n),
    "add transposed": (lambda: B.add(sq, B.permute(B.reshape(b, [c, r]), [1, 0])), 3 * ES * n),
    "contiguous(transpose)": (lambda: B.contiguous(B.permute(sq, [1, 0])), 2 * ES * n),
    "sub rowmax [N/256,256]-[N/256,1]": (lambda: B.sub(short, B.expand(B.shrink(short, [(0, n // 256), (0, 1)]), [n // 256, 256])), 2 * ES * n),
}
timing = os.environ.get("FOLD_CASES_TIME") == "1"
res = {}
for name, (fn, nbytes) in cases.items():
    fn()
    torch.cuda.synchronize()
    if not timing:
        fn()
        torch.cuda.synchronize()
        continue
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[5]
    res[name] = {"ms": round(ms, 4), "gbs": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / HBM, 3)}
if timing:
    print(json.dumps(res, indent=1))
