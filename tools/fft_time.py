#!/usr/bin/env python
"""Device time of the fft rows on one B200, CUDA-event timed through
the C ABI on the context's stream: batched power-of-two lines, one long line, a Bluestein length,
2-D, rfft. `gbs` counts the ALGORITHMIC bytes of one transform (read the input once, write the
output once) -- the floor a single fused pass would reach; every axis pass here is gather ->
transform -> scatter through a double-complex work buffer, so 1.0 is not reachable by design.
Prints one JSON object."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import raven_b200.backend as B  # noqa: E402

torch.cuda.init()
stream = torch.cuda.Stream()
ctx = B.create_context(device=0, stream=stream.cuda_stream)
rng = np.random.default_rng(0)
rows = []


def timed(fn, reps=5):
    fn()
    ctx.sync()
    best = 1e9
    with torch.cuda.stream(stream):
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
    return best


CASES = [
    ("c64", (16384, 1024), [1]), ("c64", (4096, 4096), [1]), ("c64", (256, 65536), [1]), ("c64", (1, 1 << 22), [1]),
    ("c32", (16384, 1024), [1]), ("c32", (4096, 4096), [1]),
    ("c64", (4096, 1000), [1]), ("c64", (2048, 2048), [0, 1]), ("c64", (4096, 4096), [0]),
]
if len(sys.argv) > 1:   # one case, for a profiler
    CASES = [CASES[int(sys.argv[1])]]
for dt, shape, axes in CASES:
    npdt = np.complex128 if dt == "c64" else np.complex64
    x = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(npdt)
    t = B.reshape(B.from_host(ctx, x.reshape(-1)), list(shape))
    ms = timed(lambda: B.fft(t, axes))
    nbytes = 2 * x.nbytes
    rows.append({"op": "fft", "dtype": dt, "shape": list(shape), "axes": axes, "ms": round(ms, 4),
                 "gbs": round(nbytes / ms / 1e6, 1)})
    print(rows[-1], file=sys.stderr)
for dt, shape in (("f64", (4096, 4096)), ("f32", (16384, 1024))) if len(sys.argv) == 1 else ():
    npdt = np.float64 if dt == "f64" else np.float32
    x = rng.standard_normal(shape).astype(npdt)
    t = B.reshape(B.from_host(ctx, x.reshape(-1)), list(shape))
    cdt = "c64" if dt == "f64" else "c32"
    ms = timed(lambda: B.rfft(t, cdt, [1]))
    nbytes = x.nbytes + (shape[1] // 2 + 1) * shape[0] * x.itemsize * 2
    rows.append({"op": "rfft", "dtype": dt, "shape": list(shape), "axes": [1], "ms": round(ms, 4),
                 "gbs": round(nbytes / ms / 1e6, 1)})
    print(rows[-1], file=sys.stderr)
print(json.dumps({"rows": rows}))
