#!/usr/bin/env python
"""Where bench.py's end-to-end step spends its time: each phase of the step (pinned uploads,
the seven ops, small blocking read-backs, the 1 GiB async read-back) timed alone by wall clock
with a full drain after it, then the pipelined step as bench.py runs it. One JSON object."""
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
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = B.create_context(device=0, stream=stream.cuda_stream)
pa, pb, pr = (ctx.pinned_empty(n, np.float32) for _ in range(3))
pa[:] = 1.0
pb[:] = 2.0


def drain():
    ctx.sync()
    torch.cuda.synchronize()


def wall(fn, reps=3):
    fn()
    drain()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        drain()
        ts.append(time.perf_counter() - t0)
        del r
    return round(min(ts) * 1e3, 2)


out = {}
out["upload_2x1GiB_ms"] = wall(lambda: (B.from_host(ctx, pa), B.from_host(ctx, pb)))
a, b = B.from_host(ctx, pa), B.from_host(ctx, pb)
A = B.reshape(a, [n // side, side])


def ops():
    return (B.add(a, b), B.mul(a, b), B.sin(a), B.reduce(a, "sum", [0]), B.reduce(A, "sum", [0]),
            B.reduce(A, "sum", [1]), B.argmax(a, 0))


out["seven_ops_ms"] = wall(ops)
r = ops()
drain()
out["small_readbacks_ms"] = wall(lambda: [B.to_host(x) for x in r[3:]])
out["async_readback_1GiB_ms"] = wall(lambda: B.to_host_async(r[0], pr))


def step():
    global a, b, A
    a = B.from_host(ctx, pa)
    b = B.from_host(ctx, pb)
    A = B.reshape(a, [n // side, side])
    rr = ops()
    got = [B.to_host(x) for x in rr[3:]]
    B.to_host_async(rr[0], pr)
    return got


def pipelined():
    step()
    drain()
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    drain()
    return round((time.perf_counter() - t0) / 5 * 1e3, 2)


out["pipelined_step_ms"] = pipelined()
# the same with an nvidia-smi poller running beside it (what bench.py's clock sampler does)
import subprocess  # noqa: E402
poller = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms",
                           "100", "-i", "0"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
time.sleep(0.5)
out["pipelined_step_ms_with_nvidia_smi_poller"] = pipelined()
poller.terminate()
poller.wait()
out["copy_engines"] = os.environ.get("NX_CUDA_COPY_ENGINES", "1")
print(json.dumps(out))
