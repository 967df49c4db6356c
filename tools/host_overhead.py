#!/usr/bin/env python
"""Per-op HOST cost of the Python mirror (raven_b200/backend.py over ctypes): tiny tensors so the
kernels are negligible; reports microseconds per op for a few op kinds and a cProfile breakdown.
The OCaml veneer (packages/nx-cuda) pays none of the ctypes / descriptor-building cost; this is
what bounds small-op sequences (tools/dp_step.py) when driven from Python."""
import cProfile
import io
import json
import os
import pstats
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
x = B.from_host(ctx, np.arange(1024, dtype=np.float32))
y = B.from_host(ctx, np.ones(1024, dtype=np.float32))
m = B.reshape(x, [32, 32])
N = 3000


def loop(fn):
    for _ in range(200):
        fn()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(N):
        fn()
    dt = time.perf_counter() - t0
    ctx.sync()
    return round(dt / N * 1e6, 2)


res = {
    "add": loop(lambda: B.add(x, y)),
    "add (expanded scalar operand)": loop(lambda: B.add(x, B.expand(B.reshape(y, [1024])[0:1] if False else B.shrink(y, [(0, 1)]), [1024]))),
    "sin": loop(lambda: B.sin(x)),
    "reduce sum": loop(lambda: B.reduce(m, "sum", [1])),
    "matmul 32x32": loop(lambda: B.matmul(m, m)),
    "permute (view only)": loop(lambda: B.permute(m, [1, 0])),
}
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    B.add(x, y)
pr.disable()
ctx.sync()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
print(json.dumps({"us_per_op": res}))
print(s.getvalue(), file=sys.stderr)
