#!/usr/bin/env python
"""Cholesky (and qr / triangular_solve with n right-hand sides) of large real matrices on one B200: the panel-blocked path
(nxc_linalg.cu, nxc_cholesky_blocked) against the one-CTA-per-matrix kernel it replaces there
(NX_CUDA_CHOLESKY_BLOCKED=0) and against the host LAPACK numpy links (the reference's CPU backend
calls its own unblocked loops; numpy's time is the stronger CPU bar). Host-synchronised wall time,
best of 3. Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raven_b200.backend as B  # noqa: E402

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
    for batch, n in ((1, 256), (1, 512), (1, 1024), (1, 2048), (1, 4096), (4, 1024), (16, 256), (64, 128), (256, 128), (148, 256)):
        a = rng.standard_normal((batch, n, n)).astype(npdt)
        spd = a @ np.swapaxes(a, -1, -2) / n + np.eye(n, dtype=npdt)
        ts = B.reshape(B.from_host(ctx, spd.reshape(-1)), [batch, n, n])
        r = {"dtype": dt, "batch": batch, "n": n}
        for name, env in (("blocked_ms", "1"), ("one_cta_ms", "0")):
            if env == "0" and n > 1024:
                continue  # tens of seconds: the path this one replaces
            os.environ["NX_CUDA_CHOLESKY_BLOCKED"] = env
            r[name] = t(lambda: B.cholesky(ts))
        os.environ.pop("NX_CUDA_CHOLESKY_BLOCKED")
        r["default_ms"] = t(lambda: B.cholesky(ts))
        if batch == 1 or n <= 256:
            ta = B.reshape(B.from_host(ctx, a.reshape(-1)), [batch, n, n])
            for op, fn, var in (("qr", lambda: B.qr(ta), "NX_CUDA_QR_BLOCKED"), ("trsm", lambda: B.triangular_solve(ts, ta), "NX_CUDA_TRSM_BLOCKED")):
                r[op + "_blocked_ms"] = t(fn)
                if n <= 512:
                    os.environ[var] = "0"
                    r[op + "_one_cta_ms"] = t(fn)
                    os.environ.pop(var)
        t0 = time.perf_counter()
        np.linalg.cholesky(spd)
        r["numpy_lapack_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
        rows.append(r)
        print(r, file=sys.stderr)
print(json.dumps({"rows": rows}))
