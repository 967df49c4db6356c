"""CPU checks of the workload replays (tools/tape.py, tools/gpt2_step.py) on the oracle backend:

* the tape's pull-backs are right: reverse-mode gradients of the tiny GPT-2 loss (every op class
  the 124M step uses: gather, layer norm, attention with a causal `where`, softmax, gelu, tied LM
  head, sparse cross-entropy) agree with central finite differences in float64;
* the training step is what the reference's example does: SGD lowers the loss on a fixed batch
  (train.ml's protocol), AdamW follows Vega's formulas against a numpy restatement;
* data parallelism is `mean of per-shard gradients`: two shards reduced through
  FlatBucketReducer give the gradient of the concatenated batch (jit.ml:181-190's contract)."""
import numpy as np
import pytest

from raven_b200 import dtype as D
from raven_b200 import sharded
from tests.backend_double import Ctx, OracleBackend
from tools import gpt2_step as G
from tools.tape import Tape

CFG = dict(vocab=19, n_pos=8, n_embd=8, n_layer=1, n_head=2, n_inner=16, eps=1e-5)


def _host_params(cfg, dtype, seed=0):
    rng = np.random.default_rng(seed)
    out = {}
    for k, shp in G.param_shapes(cfg).items():
        base = 1.0 if k.endswith(".g") else 0.0
        out[k] = (base + rng.standard_normal(shp) * 0.2).astype(dtype)
    return out


def _loss_and_grads(be, ctx, host, ids, targets, cfg, dt):
    P = {k: be.reshape(be.from_host(ctx, v.reshape(-1), dt), list(v.shape)) for k, v in host.items()}
    tp = Tape(be, ctx)
    tp.pos_ids = be.from_host(ctx, np.arange(ids.shape[1], dtype=np.int32))
    leaves = {k: tp.leaf(v) for k, v in P.items()}
    mask = G.causal_mask(be, ctx, ids.shape[1])
    tids = be.reshape(be.from_host(ctx, ids.reshape(-1).astype(np.int32)), list(ids.shape))
    ttg = be.from_host(ctx, targets.reshape(-1).astype(np.int32))
    loss = G.objective(tp, leaves, tids, ttg, cfg, mask)
    tp.backward(loss, be.full(ctx, loss.dtype, [], 1.0))
    return float(be.to_numpy(loss.t)), {k: be.to_numpy(tp.grad_of(v)) for k, v in leaves.items()}


def test_tape_gradients_match_finite_differences():
    be, ctx = OracleBackend(), Ctx()
    rng = np.random.default_rng(3)
    ids = rng.integers(0, CFG["vocab"], (2, 6))
    tg = rng.integers(0, CFG["vocab"], (2, 6))
    host = _host_params(CFG, np.float64)
    loss, grads = _loss_and_grads(be, ctx, host, ids, tg, CFG, D.float64)
    assert np.isfinite(loss) and abs(loss - np.log(CFG["vocab"])) < 2.0
    eps = 1e-6
    checked = 0
    for name in ("wte", "wpe", "h0.ln1.g", "h0.attn.q.w", "h0.attn.k.b", "h0.attn.v.w", "h0.attn.out.w", "h0.ln2.b",
                 "h0.fc.w", "h0.fc.b", "h0.proj.w", "ln_f.g"):
        flat = host[name].reshape(-1)
        for j in rng.choice(flat.size, size=min(3, flat.size), replace=False):
            keep = flat[j]
            flat[j] = keep + eps
            lp, _ = _loss_and_grads(be, ctx, host, ids, tg, CFG, D.float64)
            flat[j] = keep - eps
            lm, _ = _loss_and_grads(be, ctx, host, ids, tg, CFG, D.float64)
            flat[j] = keep
            fd = (lp - lm) / (2 * eps)
            an = grads[name].reshape(-1)[j]
            assert abs(fd - an) <= 1e-6 + 1e-5 * abs(fd), f"{name}[{j}]: finite difference {fd} vs tape {an}"
            checked += 1
    assert checked >= 30


def test_sgd_step_lowers_the_loss_and_adamw_follows_vega():
    be, ctx = OracleBackend(), Ctx()
    rng = np.random.default_rng(5)
    grid = rng.integers(0, CFG["vocab"], (2, 7))
    host = _host_params(CFG, np.float32, seed=1)
    tr = G.Trainer(be, ctx, CFG, 2, 6, opt="sgd", lr=0.05, host_params=host)
    tr.set_batch(grid[:, :-1], grid[:, 1:])
    losses = []
    for _ in range(5):
        tr.pre_step()
        losses.append(float(be.to_numpy(tr.step_body())))
    assert all(b < a for a, b in zip(losses, losses[1:])), losses

    tr = G.Trainer(be, ctx, CFG, 2, 6, opt="adamw", lr=1e-2, host_params=host)
    tr.set_batch(grid[:, :-1], grid[:, 1:])
    p0 = host["h0.fc.w"].astype(np.float64)
    tr.pre_step()
    _, g = _loss_and_grads(be, ctx, host, grid[:, :-1], grid[:, 1:], CFG, D.float32)
    tr.step_body()
    g = g["h0.fc.w"].astype(np.float64)
    m, n = 0.1 * g, 0.001 * g * g                                   # vega.ml:851-906, step 1
    d = (m / (1 - 0.9)) / (np.sqrt(n / (1 - 0.999)) + 1e-8)
    want = p0 - 1e-2 * (d + 0.01 * p0)
    got = be.to_numpy(tr.params["h0.fc.w"]).astype(np.float64)
    assert np.allclose(got, want, rtol=2e-4, atol=2e-6), float(np.abs(got - want).max())


class _TwoShardComm:
    """Both ranks in one process: allreduce(sum) adds the other shard's contribution."""

    def __init__(self, be, world=2):
        self.be, self.world, self.other = be, world, None

    def allreduce(self, t, op):
        assert op == "sum"
        return self.be.add(t, self.other.pop(0))


def test_data_parallel_mean_of_shard_gradients_is_the_full_batch_gradient():
    be, ctx = OracleBackend(), Ctx()
    rng = np.random.default_rng(9)
    ids = rng.integers(0, CFG["vocab"], (4, 6))
    tg = rng.integers(0, CFG["vocab"], (4, 6))
    host = _host_params(CFG, np.float64, seed=2)
    _, full = _loss_and_grads(be, ctx, host, ids, tg, CFG, D.float64)
    shard = [_loss_and_grads(be, ctx, host, ids[r * 2:(r + 1) * 2], tg[r * 2:(r + 1) * 2], CFG, D.float64)[1]
             for r in range(2)]
    comm = _TwoShardComm(be)
    red = sharded.FlatBucketReducer(comm, bucket_bytes=4096, backend=be)
    names = list(full)
    # rank 1's flat buckets, built with the same bucketing, are what rank 0's all-reduce receives
    other = sharded.FlatBucketReducer(_TwoShardComm(be), bucket_bytes=4096, backend=be)
    sent = []
    other.comm.allreduce = lambda t, op: (sent.append(t), t)[1]
    for k in names:
        other.push(k, be.reshape(be.from_host(ctx, shard[1][k].reshape(-1), D.float64), list(shard[1][k].shape)))
    other.finish()
    comm.other = sent
    n_buckets = len(sent)
    for k in names:
        red.push(k, be.reshape(be.from_host(ctx, shard[0][k].reshape(-1), D.float64), list(shard[0][k].shape)))
    avg = red.finish()
    assert n_buckets > 1 and not sent, "several buckets expected, all consumed"
    for k in names:
        assert np.allclose(be.to_numpy(avg[k]), full[k], rtol=1e-9, atol=1e-12), k
