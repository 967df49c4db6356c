#!/usr/bin/env python
"""Data-parallel training-step replay on N GPUs (one process per GPU, torchrun):
the Rune.grad-shaped MLP step of tools/mlp_grad.py (BASELINE.json configs[3]) with the
Kaun-style data-parallel exchange bolted on: parameters replicated, batch sharded on axis 0
(per-GPU batch fixed = weak scaling, as the reference's pmap shards, rune/lib/jit.ml:181-190),
after value_and_grad every gradient leaf is sum-allreduced over NCCL/NVLink and scaled by
1/world (raven_b200.sharded.allreduce_mean_), then an SGD update (mul, sub) per leaf.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dp_step.py

Rank 0 prints one JSON line: ms/step (max over ranks, CUDA events), samples/s, and the share
of the step spent in the exchange. Scaling efficiency = samples/s(N) / (N * samples/s(1))."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raven_b200.backend as B  # noqa: E402
from raven_b200 import dtype as D  # noqa: E402
from raven_b200 import sharded  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as td
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = B.create_context(device=local, stream=stream.cuda_stream)
comm = None
if world > 1:
    def exchange(idbytes):
        t = torch.tensor(list(idbytes), dtype=torch.uint8, device="cuda")
        td.broadcast(t, 0)
        return bytes(t.cpu().tolist())
    comm = sharded.NcclComm(ctx, rank, world, exchange)

BATCH, WIDTH, LAYERS = int(os.environ.get("MLP_BATCH", 8192)), int(os.environ.get("MLP_WIDTH", 4096)), 4
dt = D.of(os.environ.get("MLP_DTYPE", "bf16"))


def rand(shape, scale, seed):
    rng = np.random.default_rng(seed)
    n = int(np.prod(shape))
    blk = min(n, 1 << 22)
    t = B.from_host(ctx, (rng.standard_normal(blk) * scale).astype(np.float32))
    if blk < n:
        t = B.reshape(B.contiguous(B.expand(B.reshape(t, [1, blk]), [n // blk, blk])), [n])
    t = B.reshape(t, shape)
    return t if dt is D.float32 else B.cast(t, dt)


Ws = [rand([WIDTH, WIDTH], 1.0 / np.sqrt(WIDTH), 10 + i) for i in range(LAYERS)]   # replicated (same seed)
bs = [rand([WIDTH], 0.01, 20 + i) for i in range(LAYERS)]
x = rand([BATCH, WIDTH], 1.0, 100 + rank)                                             # this rank's batch shard
y = rand([BATCH, WIDTH], 1.0, 200 + rank)
zero, inv = B.full(ctx, dt, [], 0.0), B.full(ctx, dt, [], 1.0 / (BATCH * WIDTH))
two_inv = B.mul(inv, B.full(ctx, dt, [], 2.0))
lr = B.full(ctx, dt, [], 1e-3)


OVERLAP = os.environ.get("DP_OVERLAP", "1") == "1"


def grads(reducer=None):
    hs, pres, h = [x], [], x
    for W, b in zip(Ws, bs):
        pre = B.add(B.matmul(h, W), B.expand(B.reshape(b, [1, WIDTH]), [BATCH, WIDTH]))
        h = B.max(pre, B.expand(zero, [BATCH, WIDTH]))
        pres.append(pre)
        hs.append(h)
    diff = B.sub(h, y)
    loss = B.mul(B.reduce(B.mul(diff, diff), "sum", [0, 1]), inv)
    g = B.mul(diff, B.expand(two_inv, [BATCH, WIDTH]))
    out = []
    for li in range(LAYERS - 1, -1, -1):
        g = B.mul(g, B.cast(B.cmplt(B.expand(zero, [BATCH, WIDTH]), pres[li]), dt))
        dW, db = B.matmul(B.permute(hs[li], [1, 0]), g), B.reduce(g, "sum", [0])
        if reducer is not None:   # the bucket's exchange runs under the next layer's backward GEMMs
            dW, db = reducer.push(dW), reducer.push(db)
        out.append((li, dW, db))
        if li > 0:
            g = B.matmul(g, B.permute(Ws[li], [1, 0]))
    return loss, out


ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]


def step(timed=False):
    if timed:
        ev[0].record(stream)
    reducer = sharded.GradBucketReducer(comm) if (comm is not None and OVERLAP) else None
    loss, gs = grads(reducer)
    if timed:
        ev[1].record(stream)
    leaves = [t for _, dW, db in gs for t in (dW, db)]
    if reducer is not None:
        leaves = reducer.finish()
    elif comm is not None:
        leaves = sharded.allreduce_mean_(leaves, comm)
    if timed:
        ev[2].record(stream)
    for i, (li, _, _) in enumerate(gs):                                               # SGD
        dW, db = leaves[2 * i], leaves[2 * i + 1]
        Ws[li] = B.sub(Ws[li], B.mul(dW, B.expand(lr, [WIDTH, WIDTH])))
        bs[li] = B.sub(bs[li], B.mul(db, B.expand(lr, [WIDTH])))
    if timed:
        ev[3].record(stream)
    return loss


for _ in range(3):
    step()
if world > 1:
    td.barrier()
torch.cuda.synchronize()
import time
reps, tot, comp, exch, host = 8, 0.0, 0.0, 0.0, 0.0
for _ in range(reps):
    t0 = time.perf_counter()
    loss = step(timed=True)
    host += (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    tot += ev[0].elapsed_time(ev[3])
    comp += ev[0].elapsed_time(ev[1])
    exch += ev[1].elapsed_time(ev[2])
ms = tot / reps
if world > 1:
    t = torch.tensor([ms], device="cuda")
    td.all_reduce(t, op=td.ReduceOp.MAX)
    ms = float(t.item())
if rank == 0:
    nparams = LAYERS * (WIDTH * WIDTH + WIDTH)
    print(json.dumps({"workload": f"DP MLP step: {LAYERS}x{WIDTH}, per-GPU batch {BATCH}, {dt.name}, grad allreduce "
                                  f"of {nparams} params ({nparams * dt.itemsize / 1e6:.0f} MB) + SGD",
                      "n_gpus": world, "ms_per_step": round(ms, 3), "samples_per_s": round(world * BATCH / (ms * 1e-3), 1),
                      "grad_ms": round(comp / reps, 3), "host_issue_ms": round(host / reps, 3), "exchange_ms": round(exch / reps, 3),
                      "exchange": "bucketed allreduce on the comm stream under the backward pass" if OVERLAP
                      else "allreduce after the backward pass", "scaling": "weak"}))
if comm is not None:
    ctx.sync()
    comm.close()
if world > 1:
    td.destroy_process_group()
