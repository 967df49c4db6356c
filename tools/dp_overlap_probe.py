#!/usr/bin/env python
"""How much of the gradient all-reduce the GPT-2 data-parallel step hides behind its backward pass:
the same captured step with flat buckets of several sizes (one huge bucket = everything exposed after
the backward pass) at the reference's 4 x 64-token batch and at 8 x 1024. Run under torchrun; rank 0
prints one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import gpt2_step  # noqa: E402

import torch  # noqa: E402
import torch.distributed as td  # noqa: E402
import raven_b200.backend as B  # noqa: E402
from raven_b200 import sharded  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
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
out = {"n_gpus": world, "rows": []}
for batch, seq, dt, opt in ((4, 64, "f32", "sgd"), (8, 1024, "bf16", "adamw")):
    for label, c, mb in (("no all-reduce (each rank alone)", None, 64), ("buckets of 16 MB", comm, 16),
                         ("buckets of 64 MB", comm, 64), ("one bucket", comm, 4096)):
        if c is None and world == 1 and label != "no all-reduce (each rank alone)":
            continue
        r = gpt2_step.run(batch, seq, dt, opt, steps=5, warmup=2, bucket_mb=mb, ctx=ctx, comm=c, stream=stream)
        out["rows"].append({"config": f"{batch}x{seq} {dt} {opt}", "reducer": label, "ms_per_step": r["ms_per_step"]})
        if rank == 0:
            print(out["rows"][-1], file=sys.stderr, flush=True)
        if world == 1:
            break
if rank == 0:
    print(json.dumps(out))
