#!/usr/bin/env python
"""Host<->device copy rates of this box, through the backend's own entry points (pinned upload
engine, async read-back) and through torch's pinned copies for comparison. Prints one JSON
object. Explains bench.py's `e2e` figure: that number is bounded by these rates, not by kernels."""
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
ctx = B.create_context(device=0)
pa = ctx.pinned_empty(n, np.float32)
pr = ctx.pinned_empty(n, np.float32)
pa[:] = 1.0
out = {}


def wall(fn, reps=3):
    fn()
    ctx.sync()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ctx.sync()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


t = wall(lambda: B.from_host(ctx, pa))
out["nxc_h2d_pinned_gbs"] = round(4 * n / t / 1e9, 1)
d = B.from_host(ctx, pa)
t = wall(lambda: B.to_host_async(d, pr))
out["nxc_d2h_async_gbs"] = round(4 * n / t / 1e9, 1)


def both():
    x = B.from_host(ctx, pa)
    B.to_host_async(d, pr)
    return x


t = wall(both)
out["nxc_duplex_gbs"] = round(8 * n / t / 1e9, 1)
pageable = np.ones(n, np.float32)
t = wall(lambda: B.from_host(ctx, pageable), reps=2)
out["nxc_h2d_pageable_gbs"] = round(4 * n / t / 1e9, 1)
t = wall(lambda: B.to_host(d), reps=2)
out["nxc_d2h_pageable_gbs"] = round(4 * n / t / 1e9, 1)

tp = torch.empty(n, dtype=torch.float32).pin_memory()
tg = torch.empty(n, dtype=torch.float32, device="cuda")
t = wall(lambda: tg.copy_(tp, non_blocking=True))
out["torch_h2d_pinned_gbs"] = round(4 * n / t / 1e9, 1)
t = wall(lambda: tp.copy_(tg, non_blocking=True))
out["torch_d2h_pinned_gbs"] = round(4 * n / t / 1e9, 1)
out["host_cores"] = os.cpu_count()
print(json.dumps(out))
