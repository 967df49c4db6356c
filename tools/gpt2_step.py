#!/usr/bin/env python
"""BASELINE.json configs[4]: the Kaun GPT-2-small training step, data-parallel, replayed through
the backend as the eager primitive sequence the reference's stack emits.

What is replayed (reference: packages/kaun/examples/04-gpt2/train.ml:37-49, 108-121, 160-176 and
gpt2.ml:165-200): GPT-2 124M (12 layers, 12 heads, 768 wide, vocab 50257, LM head tied to wte),
synthetic token ids, mean sparse cross-entropy, `Rune.value_and_grad` (tools/tape.py restates its
pull-backs), then the optimizer as Vega writes it (packages/vega/lib/vega.ml:813-906: plain SGD
as the example uses, or AdamW). Data parallelism as `Rune.pmap2` defines it
(packages/rune/lib/jit.ml:181-190): parameters replicated, the batch sharded on axis 0, gradients
averaged across ranks -- here by bucketed NCCL all-reduces on the communication stream that start
as the backward pass finishes each bucket (raven_b200.sharded.FlatBucketReducer). The default
bucket (512 MB) holds all of GPT-2 small's gradient: measured at N = 2 (tools/dp_overlap_probe.py,
profiles/dp_overlap_r02_n2.json) the launch-bound 4 x 64 step is FASTER with one exposed all-reduce
(19.1 ms) than with 64 MB buckets racing the backward pass for the device (20.3 ms; 17.1 ms without
any exchange), and under the bf16 sandwich every leaf becomes final only when the casts at the top
of the step are pulled back, i.e. at the very end, so there is nothing to overlap with.

One backend call per primitive, a fresh output per call, nothing fused. The whole step is a few
thousand launches, so by default it is CAPTURED once (Context.capture) and replayed; parameters
and optimizer state live in persistent buffers that the step ends by assigning into, which is
what makes the captured step replayable. `--eager` issues every op from the host instead.

    python tools/gpt2_step.py [--batch 4 --seq 64] [--dtype f32|bf16] [--opt sgd|adamw] [--eager]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/gpt2_step.py ...

Rank 0 prints one JSON line (ms/step max over ranks, tokens/s, launches/step, loss trajectory)."""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tools.tape import Tape, V  # noqa: E402

GPT2_SMALL = dict(vocab=50257, n_pos=1024, n_embd=768, n_layer=12, n_head=12, n_inner=3072, eps=1e-5)
GPT2_TINY = dict(vocab=257, n_pos=32, n_embd=32, n_layer=2, n_head=4, n_inner=64, eps=1e-5)


def param_shapes(cfg):
    C, I = cfg["n_embd"], cfg["n_inner"]
    shapes = {"wte": [cfg["vocab"], C], "wpe": [cfg["n_pos"], C]}
    for i in range(cfg["n_layer"]):
        p = f"h{i}."
        shapes.update({p + "ln1.g": [C], p + "ln1.b": [C], p + "ln2.g": [C], p + "ln2.b": [C],
                       p + "fc.w": [C, I], p + "fc.b": [I], p + "proj.w": [I, C], p + "proj.b": [C]})
        for n in ("q", "k", "v", "out"):
            shapes.update({p + f"attn.{n}.w": [C, C], p + f"attn.{n}.b": [C]})
    shapes.update({"ln_f.g": [C], "ln_f.b": [C]})
    return shapes


def init_params_host(cfg, seed=0):
    """Random-init weights of the architecture (no checkpoints offline): N(0, 0.02) matrices, unit
    gains, zero biases -- the same on every rank (replicated)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shp in param_shapes(cfg).items():
        if name.endswith(".g"):
            out[name] = np.ones(shp, dtype=np.float32)
        elif name.endswith(".b"):
            out[name] = np.zeros(shp, dtype=np.float32)
        else:
            out[name] = (rng.standard_normal(shp) * 0.02).astype(np.float32)
    return out


def upload_params(B, ctx, host):
    return {k: B.reshape(B.from_host(ctx, v.reshape(-1)), list(v.shape)) for k, v in host.items()}


def causal_mask(B, ctx, T):
    """`tril` as a bool tensor [T, T]; broadcast to [B, H, T, T] as a view at the use site."""
    m = np.tril(np.ones((T, T), dtype=np.uint8))
    from raven_b200 import dtype as D
    return B.reshape(B.from_host(ctx, m.reshape(-1), D.bool_), [T, T])


def objective(tp: Tape, P, ids, targets_flat, cfg, mask, compute=None):
    """Gpt2.logits + Loss.softmax_cross_entropy_sparse (gpt2.ml:178-200, train.ml:108-121).
    `compute`: a 16-bit dtype for the astype sandwich (train.ml:137-143): parameters are cast at the
    top, layer norm and the attention scores run in float32 islands, logits come back up to float32."""
    B = tp.B
    from raven_b200 import dtype as D
    f32 = D.float32
    Bt, T = ids.shape
    C, H = cfg["n_embd"], cfg["n_head"]
    if compute is not None:
        P = {k: tp.cast(v, compute) for k, v in P.items()}
    low = compute is not None

    def take_rows(table: V, idx, n_idx):
        # Embedding.apply -> Nx.take ~axis:0 (frontend.ml:1417-1435): flattened indices broadcast
        # along the row as a stride-0 view, one gather, reshape
        iv = B.expand(B.reshape(idx, [n_idx, 1]), [n_idx, C])
        return tp.gather(table, iv, 0)

    def ln(x, g, b):
        if low:   # layer_norm.ml:63-68: normalise in float32, scale and shift in the compute dtype
            mu_in = tp.cast(x, f32)
            mu = tp.mean(mu_in, [-1], keepdims=True)
            xc = tp.sub(mu_in, tp.bcast(mu, mu_in.shape))
            var = tp.mean(tp.mul(xc, xc), [-1], keepdims=True)
            nrm = tp.cast(tp.div(xc, tp.bcast(tp.sqrt(tp.add_s(var, cfg["eps"])), xc.shape)), compute)
            return tp.add(tp.mul(nrm, tp.bcast(g, nrm.shape)), tp.bcast(b, nrm.shape))
        return tp.layer_norm(x, g, b, cfg["eps"])

    def attention(x, pre):
        D_ = C // H

        def heads(t):
            return tp.permute(tp.reshape(t, [Bt, T, H, D_]), [0, 2, 1, 3])
        q, k, v = (heads(tp.linear(x, P[pre + n + ".w"], P[pre + n + ".b"])) for n in ("q", "k", "v"))
        qs, ks = (tp.cast(q, f32), tp.cast(k, f32)) if low else (q, k)   # attention.ml: float32 island
        scores = tp.mul_s(tp.matmul(qs, tp.permute(ks, [0, 1, 3, 2])), 1.0 / math.sqrt(D_))
        m4 = B.expand(B.reshape(mask, [1, 1, T, T]), [Bt, H, T, T])
        scores = tp.where(m4, scores, tp.scalar_like(scores, float("-inf")))
        probs = tp.softmax(scores)
        if low:
            probs = tp.cast(probs, compute)
        merged = tp.reshape(tp.permute(tp.matmul(probs, v), [0, 2, 1, 3]), [Bt, T, C])
        return tp.linear(merged, P[pre + "out.w"], P[pre + "out.b"])

    pos = B.reshape(tp.pos_ids, [T])
    x = tp.add(tp.reshape(take_rows(P["wte"], B.reshape(ids, [Bt * T]), Bt * T), [Bt, T, C]),
               tp.bcast(tp.reshape(take_rows(P["wpe"], pos, T), [1, T, C]), [Bt, T, C]))
    for i in range(cfg["n_layer"]):
        p = f"h{i}."
        x = tp.add(x, attention(ln(x, P[p + "ln1.g"], P[p + "ln1.b"]), p + "attn."))
        h = tp.linear(ln(x, P[p + "ln2.g"], P[p + "ln2.b"]), P[p + "fc.w"], P[p + "fc.b"])
        x = tp.add(x, tp.linear(tp.gelu_approx(h), P[p + "proj.w"], P[p + "proj.b"]))
    h = ln(x, P["ln_f.g"], P["ln_f.b"])
    logits = tp.matmul(h, tp.permute(P["wte"], [1, 0]))          # tied LM head: h @ wte^T
    logits = tp.reshape(logits, [Bt * T, cfg["vocab"]])
    if low:
        logits = tp.cast(logits, f32)
    return tp.cross_entropy_sparse(logits, targets_flat)


class Trainer:
    """Persistent parameters (+ optimizer state) and one training step over them."""

    def __init__(self, B, ctx, cfg, batch, seq, opt="sgd", lr=1e-4, compute=None, comm=None, seed=0,
                 bucket_mb=512, host_params=None):
        from raven_b200 import dtype as D
        self.B, self.ctx, self.cfg, self.comm = B, ctx, cfg, comm
        self.batch, self.seq, self.opt, self.lr = batch, seq, opt, lr
        self.compute = D.of(compute) if compute else None
        self.f32 = D.float32
        self.params = upload_params(B, ctx, host_params if host_params is not None else init_params_host(cfg, seed))
        self.names = list(self.params)
        self.mask = causal_mask(B, ctx, seq)
        self.pos = B.from_host(ctx, np.arange(seq, dtype=np.int32))
        self.bucket_bytes = bucket_mb << 20
        self.step_no = 0
        if opt == "adamw":
            self.mu = {k: B.full(ctx, self.f32, list(v.shape), 0.0) for k, v in self.params.items()}
            self.nu = {k: B.full(ctx, self.f32, list(v.shape), 0.0) for k, v in self.params.items()}
            # bias corrections change every step: device scalars refreshed before each step, so a
            # captured step reads the current values
            self.c1 = B.full(ctx, self.f32, [], 1.0)
            self.c2 = B.full(ctx, self.f32, [], 1.0)
        self.ids = B.reshape(B.from_host(ctx, np.zeros(batch * seq, dtype=np.int32)), [batch, seq])
        self.targets = B.from_host(ctx, np.zeros(batch * seq, dtype=np.int32))

    def set_batch(self, ids_np, targets_np):
        """Refresh the token buffers in place (a captured step reads the same buffers)."""
        B = self.B
        B.assign(self.ids, B.reshape(B.from_host(self.ctx, ids_np.reshape(-1).astype(np.int32)), [self.batch, self.seq]))
        B.assign(self.targets, B.from_host(self.ctx, targets_np.reshape(-1).astype(np.int32)))

    def _scal(self, like, value):
        B = self.B
        s = B.full(self.ctx, like.dtype, [], value)
        nd = len(like.shape)
        return B.expand(B.reshape(s, [1] * nd), list(like.shape)) if nd else s

    def _bcast0(self, s, like):
        B = self.B
        nd = len(like.shape)
        return B.expand(B.reshape(s, [1] * nd), list(like.shape)) if nd else s

    def pre_step(self):
        """Host-side per-step scalars (Vega computes them on the host, vega.ml:861-862)."""
        self.step_no += 1
        if self.opt == "adamw":
            B = self.B
            B.assign(self.c1, B.full(self.ctx, self.f32, [], 1.0 - 0.9 ** self.step_no))
            B.assign(self.c2, B.full(self.ctx, self.f32, [], 1.0 - 0.999 ** self.step_no))

    def value_and_grad(self, reducer=None):
        B = self.B
        tp = Tape(B, self.ctx)
        tp.pos_ids = self.pos
        leaves = {k: tp.leaf(v) for k, v in self.params.items()}
        if reducer is not None:
            by_uid = {v.uid: k for k, v in leaves.items()}
            tp.on_leaf_final = lambda x, g: reducer.push(by_uid[x.uid], g)
        loss = objective(tp, leaves, self.ids, self.targets, self.cfg, self.mask, self.compute)
        tp.backward(loss, B.full(self.ctx, loss.dtype, [], 1.0))
        grads = {k: tp.grad_of(v) for k, v in leaves.items()}
        return loss.t, grads

    def update(self, grads):
        """Vega's SGD (vega.ml:818-838) or AdamW (vega.ml:851-906), leaf by leaf; the new value is
        assigned into the persistent buffer."""
        B = self.B
        for k in self.names:
            p, g = self.params[k], grads[k]
            if self.opt == "sgd":
                new = B.sub(p, B.mul(g, self._scal(g, self.lr)))
            else:
                m, n = self.mu[k], self.nu[k]
                m2 = B.add(B.mul(m, self._scal(m, 0.9)), B.mul(g, self._scal(g, 1.0 - 0.9)))
                n2 = B.add(B.mul(n, self._scal(n, 0.999)), B.mul(B.mul(g, g), self._scal(g, 1.0 - 0.999)))
                mu_hat = B.fdiv(m2, self._bcast0(self.c1, m2))
                nu_hat = B.fdiv(n2, self._bcast0(self.c2, n2))
                d = B.fdiv(mu_hat, B.add(B.sqrt(nu_hat), self._scal(nu_hat, 1e-8)))
                decayed = B.add(d, B.mul(p, self._scal(p, 0.01)))
                new = B.sub(p, B.mul(decayed, self._scal(p, self.lr)))
                B.assign(m, m2)
                B.assign(n, n2)
            B.assign(p, new)

    def step_body(self):
        """loss (pre-update, as train.ml records it) and the in-place update."""
        from raven_b200 import sharded
        reducer = None
        if self.comm is not None:
            reducer = sharded.FlatBucketReducer(self.comm, self.bucket_bytes, backend=self.B)
        loss, grads = self.value_and_grad(reducer)
        if reducer is not None:
            grads = reducer.finish()
        self.update(grads)
        return loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)     # train.ml:46-47: 4 x 64 tokens
    ap.add_argument("--seq", type=int, default=64)
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16", "f16"])
    ap.add_argument("--opt", default="sgd", choices=["sgd", "adamw"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--bucket-mb", type=int, default=512)
    args = ap.parse_args()
    print(json.dumps(run(args.batch, args.seq, args.dtype, args.opt, args.steps, args.warmup, args.eager, args.tiny,
                         args.bucket_mb)))


def run(batch, seq, dtype="f32", opt="sgd", steps=10, warmup=3, eager=False, tiny=False, bucket_mb=512,
        ctx=None, comm=None, stream=None):
    """Times the step on this process's GPU (and its peers under torchrun). Returns the record;
    only rank 0's is meaningful for printing. `ctx` / `comm` / `stream`: reuse bench.py's."""
    import torch

    import raven_b200.backend as B
    from raven_b200 import sharded
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    own = ctx is None
    if own:
        torch.cuda.set_device(local)
        if world > 1:
            import torch.distributed as td
            td.init_process_group("nccl", device_id=torch.device("cuda", local))
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        ctx = B.create_context(device=local, stream=stream.cuda_stream)
        if world > 1:
            def exchange(idbytes):
                t = torch.tensor(list(idbytes), dtype=torch.uint8, device="cuda")
                td.broadcast(t, 0)
                return bytes(t.cpu().tolist())
            comm = sharded.NcclComm(ctx, rank, world, exchange)
    if world > 1:
        import torch.distributed as td
    cfg = GPT2_TINY if tiny else GPT2_SMALL
    tr = Trainer(B, ctx, cfg, batch, seq, opt=opt, compute=None if dtype == "f32" else dtype, comm=comm,
                 bucket_mb=bucket_mb)
    rng = np.random.default_rng(100 + rank)                      # this rank's shard of the batch
    grid = rng.integers(0, cfg["vocab"], (batch, seq + 1))
    tr.set_batch(grid[:, :-1], grid[:, 1:])
    nparams = sum(int(np.prod(s)) for s in param_shapes(cfg).values())

    def sync_all():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    losses = []
    t_issue = 0.0
    graph = None
    n0 = ctx.launch_count()
    tr.pre_step()
    loss = tr.step_body()                                         # first step eager: pools, attributes, NCCL
    losses.append(float(B.to_numpy(loss)))
    per_step_launches = ctx.launch_count() - n0
    if not eager:
        tr.pre_step()
        with ctx.capture() as graph:
            loss = tr.step_body()
        graph.launch()
        losses.append(float(B.to_numpy(loss)))
        per_step_launches = graph.kernels

    def one():
        nonlocal loss, t_issue
        t0 = time.perf_counter()
        tr.pre_step()
        if graph is not None:
            graph.launch()
        else:
            loss = tr.step_body()
        t_issue += time.perf_counter() - t0

    for _ in range(warmup):
        one()
        losses.append(float(B.to_numpy(loss)))
    sync_all()
    t_issue = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        one()
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms = float(t.item())
    losses.append(float(B.to_numpy(loss)))
    rec = {"workload": f"Kaun GPT-2 {'tiny' if tiny else '124M'} training step replay: batch {batch} x {seq} tokens per GPU, "
                       f"{dtype}, {opt}, value_and_grad + update, one backend call per primitive",
           "n_gpus": world, "ms_per_step": round(ms, 3), "tokens_per_s": round(world * batch * seq / (ms * 1e-3), 1),
           "launches_per_step": int(per_step_launches), "issue": "eager" if eager else "captured step replayed (CUDA graph)",
           "host_issue_ms_per_step": round(t_issue / steps * 1e3, 3), "params": nparams,
           "grad_bytes_allreduced": nparams * 4 if world > 1 else 0, "losses": [round(x, 6) for x in losses],
           "arena_mb": round(graph.arena_bytes / 2 ** 20, 1) if graph is not None else None, "scaling": "weak"}
    if graph is not None:
        graph.close()
    if own:
        if comm is not None:
            ctx.sync()
            comm.close()
        if world > 1:
            td.destroy_process_group()
    return rec


if __name__ == "__main__":
    main()
