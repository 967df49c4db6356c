#!/usr/bin/env python
"""Why bench.py's end-to-end step costs what it costs: raw pinned copy rates of the box (each
direction alone, both at once), then the e2e step loop with 1 and 3 full-array read-backs, per-step
wall times (is it steady?), through the public API. One JSON object."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raven_b200.backend as B  # noqa: E402

n = 1 << 28
side = 1 << 14
out = {}

# ---- raw rates (torch pinned copies on two streams) ----
hp = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
dv = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(2)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def rate(fn, nbytes, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return round(nbytes / best / 1e9, 1)


def h2d():
    with torch.cuda.stream(s1):
        dv[0].copy_(hp[0], non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        hp[1].copy_(dv[1], non_blocking=True)


out["raw_h2d_gbs"] = rate(h2d, 4 * n)
out["raw_d2h_gbs"] = rate(d2h, 4 * n)
out["raw_both_total_gbs"] = rate(lambda: (h2d(), d2h()), 8 * n)
del hp, dv

# ---- the e2e loop ----
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = B.create_context(device=0, stream=stream.cuda_stream)
pa, pb = (ctx.pinned_empty(n, np.float32) for _ in range(2))
pr = [ctx.pinned_empty(n, np.float32) for _ in range(3)]
pa[:] = 1.0
pb[:] = 2.0


def run(readbacks, steps=8):
    a = b = A = None

    def step():
        nonlocal a, b, A
        a = b = A = None
        a = B.from_host(ctx, pa)
        b = B.from_host(ctx, pb)
        A = B.reshape(a, [n // side, side])
        r = (B.add(a, b), B.mul(a, b), B.sin(a))
        s = (B.reduce(a, "sum", [0]), B.reduce(A, "sum", [0]), B.reduce(A, "sum", [1]), B.argmax(a, 0))
        got = [B.to_host(x) for x in s]
        for t, dst in zip(r[:readbacks], pr):
            B.to_host_async(t, dst)
        return got

    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(round((time.perf_counter() - t0) * 1e3, 1))
    t0 = time.perf_counter()
    ctx.sync()
    ts.append(round((time.perf_counter() - t0) * 1e3, 1))
    return ts


out["per_step_wall_ms_1_readback"] = run(1)
out["per_step_wall_ms_3_readbacks"] = run(3)
out["per_step_wall_ms_0_readbacks"] = run(0)
print(json.dumps(out))
