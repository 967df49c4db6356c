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

`value`    GB/s of algorithmic bytes, inputs resident in HBM, timed with CUDA events on the
           engine's stream; whole-job aggregate, max over ranks. The step is issued the way a
           step is meant to be issued on this backend: the eager op sequence is CAPTURED once
           (Context.capture -- the public API) and replayed, one graph launch per step, so the
           host is out of the timed region (at N > 1 the exchange kernels couple the ranks: a
           host hiccup on one rank otherwise stalls all of them). `eager` reports the same K
           steps issued op by op from the host, for comparison.
`e2e`      the same step through the public API from HOST buffers, op by op: H2D of a and b
           from pinned memory and D2H of ALL THREE elementwise results and every reduction
           result inside the timed region.
`roofline` the dominant kernel (the flat elementwise add): algorithmic bytes per launch / its
           average CUDA-event duration, against MEASURED_PEAKS.json; `traffic` = DRAM bytes per
           launch measured in this run by one ncu pass over a probe process (null when ncu is
           not available).
`matmul`   bf16 8192^3 on the tcgen05 path, TFLOP/s against the measured bf16 peak (the second
           half of BASELINE.json's metric), with 64 rows of the product checked against float64.
`sharded_batch_matmul`  a batch-leading bf16 matmul sharded on the batch axis, every rank its slab.
`checks`   correctness asserted in this run: at N > 1 the sharded reductions against the closed
           form of the inputs, a planted maximum on the last rank and a NaN on rank 1 through
           the sharded argmax / max, bit-identity of the results across ranks; at every N the
           e2e read-backs against the host inputs.
`mlp_grad`, `gpt2_step`  BASELINE.json configs[3] and configs[4] (tools/mlp_step.py, tools/gpt2_step.py).
`cpu_baseline` / `--impl reference`: the reference's own C backend (oracle/_ref, compiled unmodified
           from the reference sources; the C restatement when that is absent) on the host cores, on
           the SAME step at the same 2^28 size, with the reference's own thread policy.

N > 1 (launched by torchrun, one rank per GPU): the arrays are leading-axis slabs, 2^28 elements
PER GPU (weak scaling); elementwise ops are independent, each reduction is the local kernel plus
ONE exchange-and-fold kernel over NVLink peer memory (raven_b200.sharded, nxc_dist_fold.cu). No
other data-path collective exists on this path.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Nx op HBM GB/s (elementwise/reduce) and matmul TFLOP/s vs B200 roofline"
LOG2N = int(os.environ.get("NX_BENCH_LOG2N", "28"))
CPU_LOG2N = int(os.environ.get("NX_BENCH_CPU_LOG2N", str(LOG2N)))
WORKLOAD = (f"nx elementwise+reduce step (add,mul,sin,sum,sum axis0,sum axis1,argmax), f32, "
            f"2^{LOG2N} elements per GPU")


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
    blk = 1 << min(24, log2n)
    a = HostView(np.tile(rng.uniform(-4, 4, blk).astype(np.float32), n // blk), "f32", [n])
    b = HostView(np.tile(rng.uniform(-4, 4, blk).astype(np.float32), n // blk), "f32", [n])
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
    # the same 2^28 configuration as the CUDA arm; ~0.7 s per step on 16 cores, so K is bounded
    steps = max(1, min(args.steps, 8))
    warm = max(1, min(args.warmup, 2))
    r = run_cpu(steps, warm, CPU_LOG2N)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "GB/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if CPU_LOG2N == LOG2N else
                       WORKLOAD + f"; reference arm timed on a 2^{CPU_LOG2N}-element sample",
                       "algorithmic_bytes_per_step_per_gpu": step_bytes(1 << LOG2N)},
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
# DRAM traffic of the dominant kernel, measured in this run (one ncu pass over a probe process)
# ---------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(device_index):
    """Run this process on the CPUs of the NUMA node the GPU hangs off, so that the pinned staging
    buffers allocated afterwards are first-touched on that node (local-allocation policy) and the
    e2e copies do not cross the socket interconnect: at N = 8 the host's memory system is the
    limiter. Returns what was done, for the JSON line; None when the
    topology is not visible (no sysfs entry, a single node, a VM reporting -1)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0 or not os.path.isdir("/sys/devices/system/node/node1"):
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"gpu": bdf, "node": node, "cpus": len(cpus)}
    except Exception:
        return None


def measure_traffic(log2n):
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if os.environ.get("NX_BENCH_NO_NCU") == "1" or not os.path.exists(ncu):
        return None, "ncu not available"
    probe = os.path.join(ROOT, "tools", "traffic_probe.py")
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
           "--kernel-name-base", "demangled", "-k", "regex:nxc_map_flat_kernel<KBin", "-c", "3", "--csv",
           sys.executable, probe, str(log2n)]
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240, cwd=ROOT)
    except Exception as e:
        return None, f"ncu probe failed: {e}"
    per = {}
    for ln in r.stdout.splitlines():
        if "dram__bytes_" not in ln:
            continue
        cells = [c.strip('"') for c in ln.split('","')]
        try:
            ident, name, unit, val = cells[0], cells[-3], cells[-2].lower(), float(cells[-1].replace(",", ""))
        except (ValueError, IndexError):
            continue
        mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit)
        if mult is None:
            continue
        per.setdefault(ident, 0.0)
        per[ident] += val * mult
    if not per:
        return None, "ncu produced no dram metrics: " + r.stdout[-200:].replace("\n", " ")
    vals = sorted(per.values())
    return vals[len(vals) // 2], f"ncu dram__bytes_read.sum + dram__bytes_write.sum, median of {len(vals)} launches, this run"


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
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[3] / configs[4] workloads")
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

    def timed(fn, reps):
        """reps calls of fn between two events on the engine's stream, barrier + sync on both sides,
        max over ranks; ms per call."""
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        sync_all()
        ms = e0.elapsed_time(e1)
        if dist:
            t = torch.tensor([ms], device="cuda")
            td.all_reduce(t, op=td.ReduceOp.MAX)
            ms = float(t.item())
        return ms / reps

    # ---- the timed step: captured once, replayed K times ---------------------------------------
    for _ in range(args.warmup):
        step()
    with ctx.capture() as graph:
        outs = step()
    for _ in range(args.warmup):
        graph.launch()
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launch_count()
    ms_per_step = timed(graph.launch, args.steps)
    launches = ctx.launch_count() - launches0
    value = world * step_bytes(n) / (ms_per_step * 1e-3) / 1e9
    # the same K steps issued op by op from the host (what round 1 reported as `value`)
    t_host0 = time.perf_counter()
    eager_ms = timed(step, args.steps)
    eager = {"ms_per_step": round(eager_ms, 4), "value": round(world * step_bytes(n) / (eager_ms * 1e-3) / 1e9, 1),
             "unit": "GB/s", "wall_ms_per_step": round((time.perf_counter() - t_host0) * 1e3 / args.steps, 4)}

    # ---- correctness of what was just timed ------------------------------------------------------
    checks = {}
    # the replayed outputs against the eager ones (same buffers in, same kernels)
    eager_out = step()
    graph.launch()
    for name, g_t, e_t in zip(("add", "mul", "sin", "sum", "sum_axis0", "sum_axis1", "argmax"), outs, eager_out):
        if name in ("add", "mul", "sin"):   # whole arrays, compared on the device: count of differing elements
            ok = int(B.to_numpy(B.reduce(B.cast(B.cmpne(g_t, e_t), D.int32), "sum", [0]))) == 0
        else:
            ok = np.array_equal(B.to_numpy(g_t), B.to_numpy(e_t))
        if not ok:
            raise SystemExit(f"check failed: replayed {name} differs from the eager result")
    checks["replay_equals_eager"] = True
    del eager_out
    # closed forms of the tiled inputs (float64 on the host); rank r's block is seeded by r
    reps_blk = n // blk
    blocks = [np.random.default_rng(r).uniform(-4, 4, blk).astype(np.float32) for r in range(world)]
    total = sum(float(bk.sum(dtype=np.float64)) * reps_blk for bk in blocks)
    mass = sum(float(np.abs(bk, dtype=np.float64).sum()) * reps_blk for bk in blocks)
    got_total = float(B.to_numpy(outs[3]))
    if abs(got_total - total) > 1e-5 * mass:
        raise SystemExit(f"check failed: sum over all ranks {got_total} vs closed form {total}")
    checks["sum_vs_closed_form_rel_err"] = abs(got_total - total) / mass
    # axis-0 sum: column c of the [rows, side] view collects the block elements at positions = c mod side
    col = sum(bk.astype(np.float64).reshape(-1, side).sum(axis=0) * reps_blk for bk in blocks) if blk >= side else None
    if col is not None:
        got_col = B.to_numpy(outs[4]).astype(np.float64)
        lim = 1e-5 * float(np.abs(np.concatenate(blocks), dtype=np.float64).reshape(-1, side).sum(axis=0).max()) * reps_blk
        if np.abs(got_col - col).max() > lim:
            raise SystemExit("check failed: sum over axis 0 vs closed form")
        checks["sum_axis0_vs_closed_form"] = True
    got_rows = B.to_numpy(outs[5]).astype(np.float64)
    if got_rows.shape != (world * rows_local,):
        raise SystemExit(f"check failed: sum over axis 1 has shape {got_rows.shape}")
    want_rows = np.concatenate([np.tile(bk.astype(np.float64).reshape(-1, side).sum(axis=1), reps_blk) for bk in blocks]) \
        if blk >= side else None
    if want_rows is not None and np.abs(got_rows - want_rows).max() > 1e-5 * 4.0 * side:
        raise SystemExit("check failed: sum over axis 1 vs closed form")
    checks["sum_axis1_vs_closed_form"] = want_rows is not None
    # argmax: the first occurrence of the largest element over all ranks' tiled blocks
    best = max(float(bk.max()) for bk in blocks)
    first_rank = next(r for r, bk in enumerate(blocks) if float(bk.max()) == best)
    want_idx = first_rank * n + int(np.argmax(blocks[first_rank]))
    if int(B.to_numpy(outs[6])) != want_idx:
        raise SystemExit(f"check failed: argmax {int(B.to_numpy(outs[6]))} vs {want_idx}")
    checks["argmax_vs_closed_form"] = True

    def plant(t, pos, value):
        B.assign(B.shrink(t, [(pos, pos + 1)]), B.full(ctx, D.float32, [1], value))

    if dist:
        # (1) a maximum planted in the LAST rank's slab must win with its global index
        keep = float(B.to_numpy(B.shrink(a, [(n - 77, n - 76)]))[0])
        if rank == world - 1:
            plant(a, n - 77, 1000.0)
        graph.launch()
        got = int(B.to_numpy(outs[6]))
        if got != (world - 1) * n + n - 77:
            raise SystemExit(f"check failed: planted maximum on rank {world - 1}: argmax = {got}")
        # (2) a NaN on rank 1 beats it (first NaN wins, nx_c_fold.c:93-101) and sticks in max
        keep2 = float(B.to_numpy(B.shrink(a, [(4242, 4243)]))[0])
        if rank == 1 % world:
            plant(a, 4242, float("nan"))
        graph.launch()
        got = int(B.to_numpy(outs[6]))
        mx = float(B.to_numpy(sharded.sharded_reduce(a, "max", [0], comm)))
        if got != (1 % world) * n + 4242 or not np.isnan(mx):
            raise SystemExit(f"check failed: NaN on rank 1: argmax = {got}, max = {mx}")
        plant(a, 4242, float(keep2))
        plant(a, n - 77, float(keep))
        graph.launch()
        # (3) every rank holds bit-identical reduction results
        import hashlib
        digest = hashlib.sha256(b"".join(np.ascontiguousarray(B.to_numpy(t)).tobytes() for t in outs[3:])).hexdigest()
        seen = [None] * world
        td.all_gather_object(seen, digest)
        if len(set(seen)) != 1:
            raise SystemExit(f"check failed: reduction results differ across ranks: {seen}")
        if int(B.to_numpy(outs[6])) != want_idx:
            raise SystemExit("check failed: argmax after restoring the inputs")
        checks.update({"planted_max_on_last_rank": True, "nan_on_rank1_wins_and_sticks": True,
                       "bit_identical_across_ranks": True, "p2p": bool(ctx._lib.nxc_dist_p2p_enabled(ctx.ptr))})

    # ---- roofline of the dominant kernel: the flat elementwise add ----------------------
    pk = peaks()
    add_ms = timed(lambda: B.add(a, b), 20)
    add_gbs = 12 * n / (add_ms * 1e-3) / 1e9
    per_op = {}

    def op_line(name, fn, nbytes, reps=10):
        fn()
        t_ms = timed(fn, reps)
        per_op[name] = {"ms": round(t_ms, 4), "gbs": round(nbytes / (t_ms * 1e-3) / 1e9, 1),
                        "frac": round(nbytes / (t_ms * 1e-3) / 1e9 / pk["hbm"], 4)}

    for name, fn, nbytes in [("add", lambda: B.add(a, b), 12 * n), ("mul", lambda: B.mul(a, b), 12 * n),
                             ("sin", lambda: B.sin(a), 8 * n), ("sum_all", lambda: B.reduce(a, "sum", [0]), 4 * n),
                             ("sum_axis0", lambda: B.reduce(A, "sum", [0]), 4 * n),
                             ("sum_axis1", lambda: B.reduce(A, "sum", [1]), 4 * n),
                             ("argmax", lambda: B.argmax(a, 0), 4 * n)]:
        op_line(name, fn, nbytes)
    if comm is not None:
        # the sharded forms of the four reductions (local kernel + exchange), same protocol
        for name, fn, nbytes in [("sharded_sum_all", lambda: sharded.sharded_reduce(a, "sum", [0], comm), 4 * n),
                                 ("sharded_sum_axis0", lambda: sharded.sharded_reduce(A, "sum", [0], comm), 4 * n),
                                 ("sharded_sum_axis1", lambda: sharded.sharded_reduce(A, "sum", [1], comm), 4 * n),
                                 ("sharded_argmax", lambda: sharded.sharded_argreduce(a, True, 0, rank * n, comm), 4 * n)]:
            op_line(name, fn, nbytes)
    roofline = {"bound": "hbm", "kernel": "nxc_map_flat_kernel<add,f32>", "achieved": round(add_gbs, 1),
                "peak": pk["hbm"], "peak_source": pk["src"] + " (burst copy)", "unit": "GB/s",
                "frac": round(add_gbs / pk["hbm"], 4), "traffic": None, "traffic_source": None,
                "algorithmic_bytes_per_launch": 12 * n, "per_op": per_op}

    # ---- matmul half of the metric: bf16 8192^3 on tcgen05 --------------------------------
    matmul = None
    if not args.no_matmul and rank == 0:
        try:
            M = 8192
            x = B.cast(B.reshape(B.shrink(a, [(0, M * M)]), [M, M]), D.bfloat16)
            y = B.cast(B.reshape(B.shrink(b, [(0, M * M)]), [M, M]), D.bfloat16)
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                B.matmul(x, y)
            torch.cuda.synchronize()
            k0.record(stream)
            for _ in range(10):
                z = B.matmul(x, y)
            k1.record(stream)
            torch.cuda.synchronize()
            mm_ms = k0.elapsed_time(k1) / 10
            tf = 2.0 * M ** 3 / (mm_ms * 1e-3) / 1e12
            matmul = {"workload": "bf16 8192x8192x8192, f32 accumulate in TMEM (tcgen05)", "ms": round(mm_ms, 4),
                      "value": round(tf, 1), "unit": "TFLOP/s", "peak": pk["bf16"], "frac": round(tf / pk["bf16"], 4),
                      "bound": "tensor"}
            # 64 rows of the timed product against float64: one bf16 ulp of the exact element plus the
            # f32 accumulation slack 2 K eps32 sum|a||b| (tests/harness.py assert_gemm_16bit's bound)
            rows = np.linspace(0, M - 1, 64).astype(np.int64)
            bits = lambda t: (B.to_host(t).astype(np.uint32) << 16).view(np.float32)  # noqa: E731
            xf = bits(x).reshape(M, M)[rows].astype(np.float64)
            yf = bits(y).reshape(M, M).astype(np.float64)
            zf = bits(z).reshape(M, M)[rows].astype(np.float64)
            exact = xf @ yf
            ulp = np.exp2(np.floor(np.log2(np.maximum(np.abs(exact), 1e-300))) - 7)
            bound = ulp + 2.0 * M * 2.0 ** -24 * (np.abs(xf) @ np.abs(yf))
            worst = float(np.max(np.abs(zf - exact) / bound))
            if not worst <= 1.0:
                raise SystemExit(f"check failed: bf16 8192^3 product, 64-row float64 check: {worst:.2f} x the bound")
            matmul["check"] = {"rows": 64, "worst_error_over_bound": round(worst, 3),
                               "bound": "1 bf16 ulp + 2*K*2^-24*sum|a||b| per element vs a float64 product"}
            del x, y, z, xf, yf, zf, exact
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
        except SystemExit:
            raise
        except Exception as e:  # report, never hide
            matmul = dict(matmul or {}, error=str(e))

    # ---- batch-leading matmul sharded on the batch axis: every rank multiplies
    # its own slab of the batch, no exchange; whole-job TFLOP/s, max over ranks -------------------
    batch_mm = None
    if not args.no_matmul:
        try:
            Bn, Mn = 8, 2048
            xa = B.cast(B.reshape(B.shrink(a, [(0, Bn * Mn * Mn)]), [Bn, Mn, Mn]), D.bfloat16)
            xb = B.cast(B.reshape(B.shrink(b, [(0, Bn * Mn * Mn)]), [Bn, Mn, Mn]), D.bfloat16)
            sharded.sharded_batch_matmul(xa, xb, comm)
            t_ms = timed(lambda: sharded.sharded_batch_matmul(xa, xb, comm), 10)
            tf = world * 2.0 * Bn * Mn ** 3 / (t_ms * 1e-3) / 1e12
            batch_mm = {"workload": f"bf16 [{Bn * world}, {Mn}, {Mn}] x [{Bn * world}, {Mn}, {Mn}], {Bn} batches per GPU",
                        "ms": round(t_ms, 4), "value": round(tf, 1), "unit": "TFLOP/s",
                        "frac_of_peak_per_gpu": round(tf / world / pk["bf16"], 4)}
            del xa, xb
        except Exception as e:
            batch_mm = {"error": str(e)}

    # ---- end to end: host buffers in, host results out --------------------------------------
    # the clock sampler covers the device-timed regions above and stops here
    clocks = sampler.stop() if sampler else None
    graph.close()
    del outs
    nbytes = 4 * n
    # pinned host buffers: inputs go up through the upload engine, all three elementwise results come
    # back through the download engine (nxc_d2h_async), so step i's read-backs overlap step i+1's
    # upload -- both PCIe directions busy; the small reduction results use the blocking to_host
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(torch.cuda.current_device())
    pa, pb = (ctx.pinned_empty(n, np.float32) for _ in range(2))
    pr = [ctx.pinned_empty(n, np.float32) for _ in range(3)]
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
        for r, dst in zip((r0, r1, r2), pr):              # queue behind the 3 GiB of read-backs
            B.to_host_async(r, dst)
        return got

    e2e_steps = max(2, min(args.steps, 10))
    # W untimed steps first, like the device-timed loop: the first pipelined steps grow the
    # stream-ordered pool (a read-back still owns its buffer when the next step allocates), and a
    # pool growth of 1 GiB costs 100+ ms of driver time
    for _ in range(args.warmup):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    per_step = []
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        res = e2e_step()
        per_step.append(round((time.perf_counter() - t1) * 1e3, 1))
    sync_all()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_local = e2e_s
    if dist:
        t = torch.tensor([e2e_s], device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        e2e_s = float(t.item())
    d2h = 3 * nbytes + sum(int(x.nbytes) for x in res)
    e2e = {"value": round(world * step_bytes(n) / e2e_s / 1e9, 2), "unit": "GB/s",
           "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_s * 1e3, 3),
           "steps": e2e_steps, "host_wall_ms_of_each_step": per_step,
           # what limits it: each rank moves this much over its own PCIe link per step, both directions at once
           "pcie_gbs_per_rank": {"h2d": round(2 * nbytes / e2e_local / 1e9, 1), "d2h": round(d2h / e2e_local / 1e9, 1)},
           # staging buffers placed on the GPU's own NUMA node (null: topology not visible, nothing bound)
           "numa_binding_rank0": numa}
    # the last step's read-backs are complete (sync_all above drains the device): check all three
    # against the host inputs -- add and mul bit for bit over the whole array, sin within 2 ulp on a sample
    if not np.array_equal(pr[0], pa + pb):
        raise SystemExit("e2e: read-back of add(a, b) does not match the host inputs")
    if not np.array_equal(pr[1], pa * pb):
        raise SystemExit("e2e: read-back of mul(a, b) does not match the host inputs")
    smp = slice(0, 1 << 20)
    ref_sin = np.sin(pa[smp].astype(np.float64))
    if np.abs(pr[2][smp].astype(np.float64) - ref_sin).max() > 2.5 * 2.0 ** -24:
        raise SystemExit("e2e: read-back of sin(a) is off")
    if int(res[3][0]) != want_idx:
        raise SystemExit("e2e: argmax read back differs")
    checks["e2e_readbacks_match_host"] = True
    del pa, pb, pr, a, b, A
    ctx.sync()
    os.sched_setaffinity(0, all_cpus)   # the CPU baseline below uses every core again

    # ---- BASELINE.json configs[3] and configs[4] through the same backend ----------------------
    mlp_grad = gpt2 = None
    if not args.no_extra:
        try:
            from tools import mlp_step
            if rank == 0:
                mlp_grad = mlp_step.run(ctx, stream)
        except Exception as e:
            mlp_grad = {"error": str(e)}
        try:
            from tools import gpt2_step
            gpt2 = {"reference_protocol_4x64": gpt2_step.run(4, 64, "f32", "sgd", steps=10, warmup=3, ctx=ctx, comm=comm, stream=stream),
                    "bf16_8x1024": gpt2_step.run(8, 1024, "bf16", "adamw", steps=5, warmup=2, ctx=ctx, comm=comm, stream=stream)}
        except Exception as e:
            gpt2 = {"error": str(e)}

    # ---- the reference's CPU backend on the host cores, same step, same size (rank 0) ----------
    cpu = None
    if rank == 0 and not args.no_cpu:
        r = run_cpu(3, 1, CPU_LOG2N)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu["value"] = round(cpu["value"], 3)
    if rank == 0 and world == 1:
        roofline["traffic"], roofline["traffic_source"] = measure_traffic(LOG2N)
    if dist:
        td.barrier()

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "algorithmic_bytes_per_step_per_gpu": step_bytes(n),
                           "issue": "the eager op sequence captured once (Context.capture) and replayed: one CUDA "
                                    "graph launch per step",
                           "l2": "inputs (1 GiB each) exceed the 126 MB L2; no flush between iterations",
                           "parallelism": f"leading-axis slabs x{world}; each reduction = local kernel + one "
                                          f"exchange-and-fold kernel over NVLink peer memory" if dist else "single GPU"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "eager": eager, "checks": checks,
                "roofline": roofline, "cpu_baseline": cpu, "matmul": matmul, "sharded_batch_matmul": batch_mm, "mlp_grad": mlp_grad, "gpt2_step": gpt2,
                "impl": "cuda"}
        print(json.dumps(line))
    if comm is not None:
        ctx.sync()
        comm.close()
    if dist:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
