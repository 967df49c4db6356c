#!/usr/bin/env python
"""Small / medium GEMMs: where the time goes. Per size (bf16 NN): the device time of one product
inside a replayed captured step (no host in the loop), the time per call when issued eagerly
(CUDA events over 100 calls = max(host issue, device)), and the host time of the call itself.
One JSON object."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raven_b200.backend as B  # noqa: E402
from raven_b200 import dtype as D  # noqa: E402

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = B.create_context(device=0, stream=stream.cuda_stream)
out = {}
dts = os.environ.get("MM_DTYPES", "bf16,f32").split(",")
for dt in dts:
    for n in (256, 512, 1024, 2048, 4096, 8192):
        rng = np.random.default_rng(n)
        a = B.reshape(B.from_host(ctx, rng.standard_normal(n * n).astype(np.float32)), [n, n])
        b = B.reshape(B.from_host(ctx, rng.standard_normal(n * n).astype(np.float32)), [n, n])
        if dt != "f32":
            a, b = B.cast(a, dt), B.cast(b, dt)
        reps = 100 if n <= 2048 else 20
        for _ in range(3):
            B.matmul(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(reps):
            c = B.matmul(a, b)
        e1.record(stream)
        host = (time.perf_counter() - t0) / reps
        torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) / reps
        with ctx.capture() as g:
            for _ in range(20):
                c = B.matmul(a, b)
        g.launch()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(5):
            g.launch()
        e1.record(stream)
        torch.cuda.synchronize()
        dev = e0.elapsed_time(e1) / 100
        g.close()
        fl = 2.0 * n ** 3
        out[f"{dt}_{n}"] = {"device_us": round(dev * 1e3, 2), "eager_us": round(eager * 1e3, 2), "host_call_us": round(host * 1e6, 2),
                            "device_tflops": round(fl / (dev * 1e-3) / 1e12, 1), "kernels_per_product": g.kernels // 20}
        del a, b, c
print(json.dumps(out))
