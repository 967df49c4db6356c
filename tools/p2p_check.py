#!/usr/bin/env python
"""Checks and times the peer-memory exchange (nxc_dist.cu: CUDA-IPC mailboxes over NVLink)
against torch.distributed's NCCL collectives. Launch one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/p2p_check.py
Rank 0 prints one JSON object: whether the mailboxes were mapped, bit-exactness of allgather and
of integer allreduce, closeness of float allreduce, and the latency of both paths per payload."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raven_b200.backend as B  # noqa: E402
from raven_b200 import dtype as D  # noqa: E402
from raven_b200 import sharded  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream(device=local)
torch.cuda.set_stream(stream)
ctx = B.create_context(device=local, stream=stream.cuda_stream)


def exchange(idbytes):
    t = torch.tensor(list(idbytes), dtype=torch.uint8, device="cuda")
    td.broadcast(t, 0)
    return bytes(t.cpu().tolist())


comm = sharded.NcclComm(ctx, rank, world, exchange)
out = {"world": world, "p2p": bool(ctx._lib.nxc_dist_p2p_enabled(ctx.ptr)), "checks": {}, "latency_us": {}}
rng = np.random.default_rng(100 + rank)
ok = True
for n in (1, 3, 1000, 4096, 16384, 65536):
    h = rng.standard_normal(n).astype(np.float32)
    t = B.from_host(ctx, h)
    g = B.to_host(comm.allgather(t)).reshape(world, n)
    ref = torch.empty(world, n, device="cuda")
    td.all_gather_into_tensor(ref, torch.from_numpy(h).cuda())
    same = bool(np.array_equal(g, ref.cpu().numpy()))
    hi = rng.integers(-1000, 1000, n).astype(np.int32)
    ri = B.to_host(comm.allreduce(B.from_host(ctx, hi), "sum"))
    refi = torch.from_numpy(hi).cuda()
    td.all_reduce(refi)
    same_i = bool(np.array_equal(ri, refi.cpu().numpy()))
    rf = B.to_host(comm.allreduce(B.from_host(ctx, h), "sum"))
    reff = torch.from_numpy(h).cuda()
    td.all_reduce(reff)
    close = bool(np.allclose(rf, reff.cpu().numpy(), rtol=1e-5, atol=1e-5))
    rm = B.to_host(comm.allreduce(B.from_host(ctx, hi), "max"))
    refm = torch.from_numpy(hi).cuda()
    td.all_reduce(refm, op=td.ReduceOp.MAX)
    same_m = bool(np.array_equal(rm, refm.cpu().numpy()))
    out["checks"][str(n)] = {"allgather_exact": same, "allreduce_i32_exact": same_i, "allreduce_f32_close": close,
                             "allreduce_max_exact": same_m}
    ok = ok and same and same_i and close and same_m


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    td.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for n in (1, 16384, 65536):
    t = B.from_host(ctx, rng.standard_normal(n).astype(np.float32))
    tt = torch.randn(n, device="cuda")
    gg = torch.empty(world * n, device="cuda")
    out["latency_us"][str(4 * n) + "B"] = {
        "nxc_allgather": round(timed(lambda: comm.allgather(t)), 2),
        "nxc_allreduce": round(timed(lambda: comm.allreduce(t, "sum")), 2),
        "torch_nccl_allgather": round(timed(lambda: td.all_gather_into_tensor(gg, tt)), 2),
        "torch_nccl_allreduce": round(timed(lambda: td.all_reduce(tt)), 2),
    }
out["ok"] = ok
flag = torch.tensor([1 if ok else 0], device="cuda")
td.all_reduce(flag, op=td.ReduceOp.MIN)
out["ok_all_ranks"] = bool(flag.item())
ctx.sync()
if rank == 0:
    print(json.dumps(out))
comm.close()
td.destroy_process_group()
