#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its configuration, on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], the point its targets are quoted on): the Nx
elementwise + axis-reduction hot path over f32 arrays of 2^28 elements per GPU.
One STEP is one pass of that path over one batch:

    add(a,b)  mul(a,b)  sin(a)  sum(a)  sum(A,axis0)  sum(A,axis1)  argmax(a)

with a, b of 2^28 f32 (1 GiB each, far larger than the 126 MB L2, so no flush is
needed between iterations) and A = a viewed as [2^14, 2^14]. Algorithmic bytes
per step: 12N + 12N + 8N + 4N + 4N + 4N + 4N = 48 N bytes.

`value`    GB/s of algorithmic bytes, inputs resident in HBM, timed with CUDA
           events on the engine's stream; whole-job aggregate, max over ranks.
`e2e`      the same step through the public API from HOST buffers: H2D of a and b
           from pinned memory and D2H of the add result and every reduction result
           inside the timed region.
`roofline` the dominant kernel (the flat elementwise add): algorithmic bytes per
           launch / its average CUDA-event duration, against MEASURED_PEAKS.json.
`matmul`   bf16 8192^3 on the tcgen05 path, TFLOP/s against the measured bf16 peak
           (the second half of BASELINE.json's metric).
`cpu_baseline` / `--impl reference`: the reference's own C backend (oracle/_ref,
           compiled unmodified from the reference sources; the C restatement when
           that is absent) on the host cores, on a bounded 2^24-element sample of
           the same step, with the reference's own thread policy.

N > 1 (launched by torchrun, one rank per GPU): the arrays are leading-axis slabs,
2^28 elements PER GPU (weak scaling); elementwise ops are independent, the
reductions are local reduce + NCCL allreduce / allgather of the tiny partials
(raven_b200.sharded). No other data-path collective exists on this path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Nx op HBM GB/s (elementwise/reduce) and matmul TFLOP/s vs B200 roofline"
LOG2N = int(os.environ.get("NX_BENCH_LOG2N", "28"))
CPU_LOG2N = int(os.environ.get("NX_BENCH_CPU_LOG2N", "24"))


def step_bytes(n):
    return 48 * n


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "bf16": float(p["bf16_tflops"]),
                "bf16_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "src": "measured"}
    except Exception:
        return {"hbm": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


# ---------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's C backend on the host cores
# ---------------------------------------------------------------------------------------
def cpu_step_runner(log2n):
    import numpy as np
    from oracle import ref, nxo
    from oracle.hostview import HostView
    use_ref = ref.available()
    m = ref if use_ref else nxo
    n = 1 << log2n
    rng = np.random.default_rng(0)
    a = HostView(rng.uniform(-4, 4, n).astype(np.float32), "f32", [n])
    b = HostView(rng.uniform(-4, 4, n).astype(np.float32), "f32", [n])
    side = 1 << (log2n // 2)
    A = HostView(a.storage, "f32", [n // side, side])

    def step():
        m.binary("add", a, b)
        m.binary("mul", a, b)
        m.unary("sin", a)
        m.reduce("sum", a, [0])
        m.reduce("sum", A, [0])
        m.reduce("sum", A, [1])
        m.argreduce("argmax", a, 0)

    return step, ("reference" if use_ref else "port"), n


def run_cpu(steps, warmup, log2n):
    step, kind, n = cpu_step_runner(log2n)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    cores = os.cpu_count() or 1
    return {"value": step_bytes(n) / dt / 1e9, "unit": "GB/s", "cores": min(cores, 64), "kind": kind,
            "sample": f"the same 7-op step on f32 arrays of 2^{log2n} elements, {steps} steps after {warmup} warm-up, "
                      f"reference thread policy (nx_c_engine.c:486-520) over {min(cores, 64)} host cores",
            "ms_per_step": dt * 1e3}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    r = run_cpu(steps, max(1, min(args.warmup, 3)), CPU_LOG2N)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "GB/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"nx elementwise+reduce step (add,mul,sin,sum,sum axis0,sum axis1,argmax), f32, "
                                   f"2^{LOG2N} elements per GPU; reference arm timed on a 2^{CPU_LOG2N}-element sample"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
# clocks sampling
# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(device)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ---------------------------------------------------------------------------------------
# the CUDA arm
# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda")
    ap.add_argument("--no-matmul", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch

    import raven_b200.backend as B
    from raven_b200 import dtype as D
    from raven_b200 import sharded

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = world > 1
    torch.cuda.set_device(local)
    if dist:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a real (non-legacy-default) stream shared by torch's events and the engine's launches
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = B.create_context(device=local, stream=stream.cuda_stream)
    assert ctx.stream() == stream.cuda_stream
    comm = None
    if dist:
        def exchange(idbytes):
            t = torch.tensor(list(idbytes), dtype=torch.uint8, device="cuda")
            td.broadcast(t, 0)
            return bytes(t.cpu().tolist())
        comm = sharded.NcclComm(ctx, rank, world, exchange)

    n = 1 << LOG2N
    side = 1 << (LOG2N // 2)
    rows_local = n // side
    rng = np.random.default_rng(rank)

    # synthetic inputs, resident in HBM. Built from a 2^24-element seeded host block
    # tiled on the device (H2D stays out of the timed region; frontend.ml:342-358).
    blk = 1 << min(24, LOG2N)
    ha = rng.uniform(-4, 4, blk).astype(np.float32)
    hb = rng.uniform(-4, 4, blk).astype(np.float32)

    def tiled(h):
        t = B.from_host(ctx, h)
        if blk == n:
            return t
        return B.contiguous(B.reshape(B.expand(B.reshape(t, [1, blk]), [n // blk, blk]), [n // blk, blk]))

    a = B.reshape(tiled(ha), [n])
    b = B.reshape(tiled(hb), [n])
    A = B.reshape(a, [rows_local, side])
    lo = rank * rows_local  # global row / element offsets of this rank's slab

    def step():
        r0 = B.add(a, b)
        r1 = B.mul(a, b)
        r2 = B.sin(a)
        if comm is None:
            s0 = B.reduce(a, "sum", [0])
            s1 = B.reduce(A, "sum", [0])
            s2 = B.reduce(A, "sum", [1])
            am = B.argmax(a, 0)
        else:
            s0 = sharded.sharded_reduce(a, "sum", [0], comm)
            s1 = sharded.sharded_reduce(A, "sum", [0], comm)
            s2 = sharded.sharded_reduce(A, "sum", [1], comm)
            am = sharded.sharded_argreduce(a, True, 0, rank * n, comm)
        return r0, r1, r2, s0, s1, s2, am

    def sync_all():
        if dist:
            td.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    if dist:
        t = torch.tensor([ms], device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * step_bytes(n) / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel: the flat elementwise add ----------------------
    pk = peaks()
    reps = 20
    out = B.add(a, b)
    sync_all()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    for _ in range(reps):
        out = B.add(a, b)
    k1.record(stream)
    torch.cuda.synchronize()
    add_ms = k0.elapsed_time(k1) / reps
    add_gbs = 12 * n / (add_ms * 1e-3) / 1e9
    per_op = {}
    for name, fn, nbytes in [("add", lambda: B.add(a, b), 12 * n), ("mul", lambda: B.mul(a, b), 12 * n),
                             ("sin", lambda: B.sin(a), 8 * n), ("sum_all", lambda: B.reduce(a, "sum", [0]), 4 * n),
                             ("sum_axis0", lambda: B.reduce(A, "sum", [0]), 4 * n),
                             ("sum_axis1", lambda: B.reduce(A, "sum", [1]), 4 * n),
                             ("argmax", lambda: B.argmax(a, 0), 4 * n)]:
        fn()
        torch.cuda.synchronize()
        k0.record(stream)
        for _ in range(10):
            fn()
        k1.record(stream)
        torch.cuda.synchronize()
        t_ms = k0.elapsed_time(k1) / 10
        per_op[name] = {"ms": round(t_ms, 4), "gbs": round(nbytes / (t_ms * 1e-3) / 1e9, 1),
                        "frac": round(nbytes / (t_ms * 1e-3) / 1e9 / pk["hbm"], 4)}
    if comm is not None:
        # the sharded forms of the four reductions (local kernel + exchange), same protocol
        for name, fn, nbytes in [("sharded_sum_all", lambda: sharded.sharded_reduce(a, "sum", [0], comm), 4 * n),
                                 ("sharded_sum_axis0", lambda: sharded.sharded_reduce(A, "sum", [0], comm), 4 * n),
                                 ("sharded_sum_axis1", lambda: sharded.sharded_reduce(A, "sum", [1], comm), 4 * n),
                                 ("sharded_argmax", lambda: sharded.sharded_argreduce(a, True, 0, rank * n, comm), 4 * n)]:
            fn()
            sync_all()
            k0.record(stream)
            for _ in range(10):
                fn()
            k1.record(stream)
            torch.cuda.synchronize()
            t_ms = k0.elapsed_time(k1) / 10
            per_op[name] = {"ms": round(t_ms, 4), "gbs": round(nbytes / (t_ms * 1e-3) / 1e9, 1),
                            "frac": round(nbytes / (t_ms * 1e-3) / 1e9 / pk["hbm"], 4)}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_add.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "nxc_map_flat_kernel<add,f32>", "achieved": round(add_gbs, 1),
                "peak": pk["hbm"], "peak_source": pk["src"] + " (burst copy)", "unit": "GB/s",
                "frac": round(add_gbs / pk["hbm"], 4), "traffic": traffic,
                "algorithmic_bytes_per_launch": 12 * n, "per_op": per_op}

    # ---- matmul half of the metric: bf16 8192^3 on tcgen05 --------------------------------
    matmul = None
    if not args.no_matmul and rank == 0:
        try:
            M = 8192
            x = B.cast(B.reshape(B.shrink(a, [(0, M * M)]), [M, M]), D.bfloat16)
            y = B.cast(B.reshape(B.shrink(b, [(0, M * M)]), [M, M]), D.bfloat16)
            for _ in range(3):
                B.matmul(x, y)
            torch.cuda.synchronize()
            k0.record(stream)
            for _ in range(10):
                B.matmul(x, y)
            k1.record(stream)
            torch.cuda.synchronize()
            mm_ms = k0.elapsed_time(k1) / 10
            tf = 2.0 * M ** 3 / (mm_ms * 1e-3) / 1e12
            matmul = {"workload": "bf16 8192x8192x8192, f32 accumulate in TMEM (tcgen05)", "ms": round(mm_ms, 4),
                      "value": round(tf, 1), "unit": "TFLOP/s", "peak": pk["bf16"], "frac": round(tf / pk["bf16"], 4),
                      "bound": "tensor"}
            del x, y
            # f32 operands, default mode: f32-class accuracy as 3xTF32 on the same tensor-core kernel
            # (two split passes + one tf32 GEMM over a tripled K axis per product, all inside the timing)
            xf = B.reshape(B.shrink(a, [(0, M * M)]), [M, M])
            yf = B.reshape(B.shrink(b, [(0, M * M)]), [M, M])
            for _ in range(2):
                B.matmul(xf, yf)
            torch.cuda.synchronize()
            k0.record(stream)
            for _ in range(4):
                B.matmul(xf, yf)
            k1.record(stream)
            torch.cuda.synchronize()
            f_ms = k0.elapsed_time(k1) / 4
            matmul["f32"] = {"workload": "f32 8192x8192x8192, default mode (3xTF32 on tcgen05, f32-class accuracy)",
                             "ms": round(f_ms, 4), "value": round(2.0 * M ** 3 / (f_ms * 1e-3) / 1e12, 1), "unit": "TFLOP/s"}
            del xf, yf
        except Exception as e:  # report, never hide
            matmul = dict(matmul or {}, error=str(e))

    # ---- end to end: host buffers in, host results out --------------------------------------
    # the clock sampler covers the device-timed regions above and stops here
    clocks = sampler.stop() if sampler else None
    del out
    nbytes = 4 * n
    # pinned host buffers: inputs go up through the upload engine, the elementwise result comes
    # back through the download engine (nxc_d2h_async), so step i's read-back overlaps step i+1's
    # upload -- both PCIe directions busy; the small reduction results use the blocking to_host
    pa, pb, pr = (ctx.pinned_empty(n, np.float32) for _ in range(3))
    pa[:] = np.tile(ha, n // blk)
    pb[:] = np.tile(hb, n // blk)

    def e2e_step():
        nonlocal a, b, A
        a = b = A = None   # release the previous step's inputs before allocating this step's
        a = B.from_host(ctx, pa)
        b = B.from_host(ctx, pb)
        A = B.reshape(a, [rows_local, side])
        r0, r1, r2, s0, s1, s2, am = step()
        got = [B.to_host(x) for x in (s0, s1, s2, am)]   # small, blocking: first, so they do not
        B.to_host_async(r0, pr)                            # queue behind the 1 GiB read-back
        return got

    e2e_steps = max(2, min(args.steps, 5))
    # W untimed steps first, like the device-timed loop: the first pipelined steps grow the
    # stream-ordered pool (a read-back still owns its buffer when the next step allocates), and a
    # pool growth of 1 GiB costs 100+ ms of driver time -- measured 46 ms/step once warm against
    # 75-185 ms when those growths landed in a 5-step timed region after a single warm-up step
    for _ in range(args.warmup):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = e2e_step()
    sync_all()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist:
        t = torch.tensor([e2e_s], device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        e2e_s = float(t.item())
    d2h = nbytes + sum(int(x.nbytes) for x in res)
    if os.environ.get("NX_BENCH_E2E_DEBUG") == "1" and rank == 0:
        # where one step's wall time goes, each phase drained before the next starts
        def _t(fn):
            t1 = time.perf_counter()
            r = fn()
            ctx.sync()
            torch.cuda.synchronize()
            return r, round((time.perf_counter() - t1) * 1e3, 2)
        (ua, ub), t_up = _t(lambda: (B.from_host(ctx, pa), B.from_host(ctx, pb)))
        a, b = ua, ub
        A = B.reshape(a, [rows_local, side])
        outs, t_ops = _t(step)
        _, t_small = _t(lambda: [B.to_host(x) for x in outs[3:]])
        _, t_back = _t(lambda: B.to_host_async(outs[0], pr))
        sys.stderr.write(f"e2e debug: upload {t_up} ms, ops {t_ops} ms, small read-backs {t_small} ms, "
                         f"1 GiB read-back {t_back} ms\n")
        each = {}
        for nm, fn in (("add", lambda: B.add(a, b)), ("mul", lambda: B.mul(a, b)), ("sin", lambda: B.sin(a)),
                       ("sum", lambda: B.reduce(a, "sum", [0])), ("sum0", lambda: B.reduce(A, "sum", [0])),
                       ("sum1", lambda: B.reduce(A, "sum", [1])), ("argmax", lambda: B.argmax(a, 0)),
                       ("add again", lambda: B.add(a, b)), ("step again", step)):
            t1 = time.perf_counter()
            r = fn()
            t_call = time.perf_counter() - t1
            ctx.sync()
            torch.cuda.synchronize()
            each[nm] = (round(t_call * 1e3, 2), round((time.perf_counter() - t1) * 1e3, 2))
            del r
        sys.stderr.write(f"e2e debug per op (call ms, call+drain ms): {each}\n")
    e2e = {"value": round(world * step_bytes(n) / e2e_s / 1e9, 2), "unit": "GB/s",
           "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_s * 1e3, 3),
           "steps": e2e_steps}
    # the last step's read-back is complete (sync_all above drains the device): check it
    if not np.array_equal(pr[:1024], (pa[:1024] + pb[:1024])):
        raise SystemExit("e2e: read-back of add(a, b) does not match the host inputs")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = run_cpu(8, 2, CPU_LOG2N)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu["value"] = round(cpu["value"], 3)

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"nx elementwise+reduce step (add,mul,sin,sum,sum axis0,sum axis1,argmax), f32, "
                                       f"2^{LOG2N} elements per GPU",
                           "algorithmic_bytes_per_step_per_gpu": step_bytes(n),
                           "l2": "inputs (1 GiB each) exceed the 126 MB L2; no flush between iterations",
                           "parallelism": f"leading-axis slabs x{world}, NCCL allreduce/allgather of reduction partials"
                           if dist else "single GPU"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "matmul": matmul, "impl": "cuda"}
        print(json.dumps(line))
    if comm is not None:
        ctx.sync()
        comm.close()
    if dist:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
