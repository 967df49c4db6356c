#!/usr/bin/env python
"""BASELINE.json configs[1]/[2] sweeps on one B200: elementwise + axis reductions over
2^20..2^30 elements, f32/f64/i32, contiguous / transposed / row- and column-broadcast /
sliced / flipped views; matmul 512^3..16384^3 bf16 / f16 / tf32 / f32-exact, NN/NT/TN.
Prints one JSON object; `python tools/sweep.py > profiles/sweep_rNN.json`.
GB/s are ALGORITHMIC bytes / CUDA-event time (median of 10 after 3
warm-ups; inputs > L2 or rotated), frac = / MEASURED_PEAKS.json hbm_gbs."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raven_b200.backend as B  # noqa: E402
from raven_b200 import dtype as D  # noqa: E402

pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
HBM, BF16 = pk["hbm_gbs"], pk["bf16_tflops"]
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = B.create_context(device=0, stream=stream.cuda_stream)


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def rand(dt, n):
    blk = min(n, 1 << 22)
    rng = np.random.default_rng(0)
    h = rng.uniform(-4, 4, blk).astype(dt.np) if dt.cls == "float" else rng.integers(-1000, 1000, blk).astype(dt.np)
    t = B.from_host(ctx, h)
    if blk < n:
        t = B.reshape(B.contiguous(B.expand(B.reshape(t, [1, blk]), [n // blk, blk])), [n])
    return t


out = {"elementwise": [], "reduce": [], "matmul": [], "peaks": {"hbm_gbs": HBM, "bf16_tflops": BF16}}
sizes = [int(x) for x in os.environ.get("SWEEP_LOG2", "20,22,24,26,28,30").split(",")]
for dt in (D.float32, D.float64, D.int32):
    es = dt.itemsize
    for lg in sizes:
        n = 1 << lg
        if n * es * 3 > 60e9:
            continue
        r = 1 << (lg // 2)
        c = n // r
        a = rand(dt, n)
        b = rand(dt, n)
        A, Bm = B.reshape(a, [r, c]), B.reshape(b, [r, c])
        cases = {
            "add contiguous": (lambda: B.add(a, b), 3 * n * es),
            "add transposed rhs": (lambda: B.add(A, B.permute(B.reshape(b, [c, r]), [1, 0])), 3 * n * es),
            "add row-broadcast [R,C]+[1,C]": (lambda: B.add(A, B.expand(B.shrink(Bm, [(0, 1), (0, c)]), [r, c])), (2 * n + c) * es),
            "add col-broadcast [R,C]+[R,1]": (lambda: B.add(A, B.expand(B.shrink(Bm, [(0, r), (0, 1)]), [r, c])), (2 * n + r) * es),
            "add sliced rows": (lambda: B.add(B.shrink(A, [(0, r), (8, c - 8)]), B.shrink(Bm, [(0, r), (8, c - 8)])), 3 * r * (c - 16) * es),
            "add flipped": (lambda: B.add(A, B.flip(Bm, [True, True])), 3 * n * es),
            "mul scalar": (lambda: B.mul(a, B.expand(B.full(ctx, dt, [], 3), [n])), 2 * n * es),
            "contiguous(transpose)": (lambda: B.contiguous(B.permute(A, [1, 0])), 2 * n * es),
            "cmplt": (lambda: B.cmplt(a, b), n * (2 * es + 1)),
            "where": (lambda: B.where(B.cmplt(a, b), a, b), None),
            "cast->f32" if dt is not D.float32 else "cast->bf16":
                (lambda: B.cast(a, D.float32 if dt is not D.float32 else D.bfloat16),
                 n * (es + (4 if dt is not D.float32 else 2))),
        }
        if dt.cls == "float":
            cases["sin"] = (lambda: B.sin(a), 2 * n * es)
            cases["exp"] = (lambda: B.exp(a), 2 * n * es)
        for name, (fn, nbytes) in cases.items():
            if nbytes is None:
                continue
            ms = timeit(fn)
            g = nbytes / (ms * 1e-3) / 1e9
            out["elementwise"].append({"dtype": dt.name, "log2n": lg, "case": name, "ms": round(ms, 4),
                                       "gbs": round(g, 1), "frac": round(g / HBM, 3)})
        shapes = {"sqrt x sqrt": (r, c), "N/256 x 256": (n // 256, 256), "256 x N/256": (256, n // 256)}
        for sname, (rr, cc) in shapes.items():
            M2 = B.reshape(a, [rr, cc])
            for op in ("sum", "max"):
                for axes, aname in (([1], "inner"), ([0], "outer"), ([0, 1], "all")):
                    ms = timeit(lambda: B.reduce(M2, op, axes))
                    g = n * es / (ms * 1e-3) / 1e9
                    out["reduce"].append({"dtype": dt.name, "log2n": lg, "shape": sname, "op": op, "axes": aname,
                                          "ms": round(ms, 4), "gbs": round(g, 1), "frac": round(g / HBM, 3)})
            for axis, aname in ((1, "inner"), (0, "outer")):
                ms = timeit(lambda: B.argmax(M2, axis))
                g = n * es / (ms * 1e-3) / 1e9
                out["reduce"].append({"dtype": dt.name, "log2n": lg, "shape": sname, "op": "argmax", "axes": aname,
                                      "ms": round(ms, 4), "gbs": round(g, 1), "frac": round(g / HBM, 3)})
        del a, b, A, Bm

if os.environ.get("SWEEP_MATMUL", "1") == "1":
    for dtn, mode in (("bf16", None), ("f16", None), ("f32", "tf32"), ("f32", "f32"), ("f32", "ieee")):
        dt = D.of(dtn)
        ctx.set_matmul_mode(mode or "f32")
        for M in (512, 1024, 2048, 4096, 8192, 16384):
            if mode == "ieee" and M > 8192:
                continue
            src = rand(D.float32, M * M)
            x = B.reshape(src if dt is D.float32 else B.cast(src, dt), [M, M])
            y = B.reshape(B.copy(x), [M, M])
            for lay, (p, q) in {"NN": (x, y), "NT": (x, B.permute(y, [1, 0])), "TN": (B.permute(x, [1, 0]), y)}.items():
                ms = timeit(lambda: B.matmul(p, q), reps=5 if M >= 8192 else 10)
                tf = 2.0 * M ** 3 / (ms * 1e-3) / 1e12
                out["matmul"].append({"dtype": dtn + ("/" + mode if mode else ""), "M": M, "layout": lay,
                                      "ms": round(ms, 4), "tflops": round(tf, 1), "frac_bf16_peak": round(tf / BF16, 3)})
            del src, x, y
    # batched [B,M,K]x[B,K,N] with B*M = 8192
    ctx.set_matmul_mode("f32")
    for M in (512, 1024):
        bsz = 8192 // M
        src = B.cast(rand(D.float32, bsz * M * M), D.bfloat16)
        x = B.reshape(src, [bsz, M, M])
        ms = timeit(lambda: B.matmul(x, x))
        tf = 2.0 * bsz * M ** 3 / (ms * 1e-3) / 1e12
        out["matmul"].append({"dtype": "bf16", "M": M, "layout": f"batched x{bsz}", "ms": round(ms, 4),
                              "tflops": round(tf, 1), "frac_bf16_peak": round(tf / BF16, 3)})
print(json.dumps(out))
