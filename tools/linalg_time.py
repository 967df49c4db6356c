#!/usr/bin/env python
"""Wall time of the linalg rows on one B200, host-synchronised calls
(these ops read a status word back, like the reference raises synchronously): batched small and
single medium matrices, f32 and f64. One CTA per matrix: these kernels are sized for the batched
small / medium factorizations ML code issues, not for one huge matrix. Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raven_b200.backend as B  # noqa: E402
from raven_b200 import dtype as D  # noqa: E402

ctx = B.create_context(device=0)
rng = np.random.default_rng(0)
rows = []


def t(fn, reps=3):
    fn()
    ctx.sync()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ctx.sync()
        best = min(best, time.perf_counter() - t0)
    return round(best * 1e3, 3)


for dt in ("f32", "f64"):
    npdt = np.float32 if dt == "f32" else np.float64
    for batch, n in ((256, 32), (64, 128), (1, 512)):
        a = rng.standard_normal((batch, n, n)).astype(npdt)
        spd = a @ np.swapaxes(a, -1, -2) + n * np.eye(n, dtype=npdt)
        ta = B.reshape(B.from_host(ctx, a.reshape(-1)), [batch, n, n])
        ts = B.reshape(B.from_host(ctx, spd.reshape(-1)), [batch, n, n])
        r = {"dtype": dt, "batch": batch, "n": n,
             "cholesky_ms": t(lambda: B.cholesky(ts)), "qr_ms": t(lambda: B.qr(ta)),
             "eigh_ms": t(lambda: B.eigh(ts)), "svd_ms": t(lambda: B.svd(ta)),
             "eig_ms": t(lambda: B.eig(ta)), "eigvals_ms": t(lambda: B.eigvals(ta))}
        rows.append(r)
        print(r, file=sys.stderr)
print(json.dumps({"rows": rows}))
