#!/usr/bin/env python
"""Which backend calls of the MLP training step block the host? Wraps every public op of
raven_b200.backend with a wall-clock timer and replays tools/dp_step.py's step at N=1."""
import os
import runpy
import sys
import time
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raven_b200.backend as B  # noqa: E402

acc = defaultdict(lambda: [0, 0.0, 0.0])
ON = [False]


def wrap(name, fn):
    def f(*a, **k):
        if not ON[0]:
            return fn(*a, **k)
        t0 = time.perf_counter()
        r = fn(*a, **k)
        dt = (time.perf_counter() - t0) * 1e6
        e = acc[name]
        e[0] += 1; e[1] += dt; e[2] = max(e[2], dt)
        return r
    return f


for n in ("add", "sub", "mul", "max", "matmul", "reduce", "cast", "cmplt", "expand", "reshape", "permute", "full"):
    setattr(B, n, wrap(n, getattr(B, n)))
os.environ["DP_PROBE"] = "1"
g = runpy.run_path(os.path.join(ROOT, "tools", "dp_step.py"), run_name="__probe__")
import torch  # noqa: E402
step = g["step"]
for _ in range(3):
    step()
torch.cuda.synchronize()
ON[0] = True
for _ in range(5):
    t0 = time.perf_counter()
    step()
    host = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    print(f"host issue {host:.3f} ms")
for n, (c, tot, mx) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:10s} calls {c:4d}  mean {tot / c:8.1f} us  max {mx:8.1f} us  total {tot / 1e3:7.3f} ms")
