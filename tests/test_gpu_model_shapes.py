"""BASELINE.json configs[3] and [4] as PARITY cases: the primitive sequences Rune / Kaun issue for a
4096-wide MLP layer and for a GPT-2-small block (kaun/examples/04-gpt2/train.ml:37-49: vocab 50257,
n_embd 768, 12 heads of 64, n_inner 3072), at those widths and a token count the oracle finishes in
seconds (2 x 128 tokens), every op checked against the reference binary on the same bytes:
embedding = gather with a stride-0 index (frontend.ml:1410-1435) and its backward = scatter Add
(reverse.ml:521-526), layer norm and softmax as the frontend composes them from reduce / broadcast
map ops, batched attention products over transposed VIEWS, the tied-embedding logits product, the
cross-entropy pieces. Tolerances: bit-exact for indices / integer / select work, 2 ulp elementwise,
1e-5 relative for reductions, f32 matmul 1e-5 * sqrt(k) of the largest output."""
import numpy as np
import pytest

import raven_b200.backend as B
from tests import harness as H

pytestmark = pytest.mark.gpu

V, D, HEADS, HD, INNER = 50257, 768, 12, 64, 3072
BT, T = 2, 128
N = BT * T


def hv(a, dt="f32"):
    return H.HostView.from_array(np.ascontiguousarray(a), dt)


def close(got, want, rel, what):
    H.assert_close("f32", got, want, rel=rel, abs_=rel * float(np.abs(want).max() or 1.0), what=what)


def test_gpt2_embedding_gather_and_scatter_add(ctx, oracle):
    rng = np.random.default_rng(31)
    wte = hv(rng.standard_normal((V, D)).astype(np.float32) * 0.02)
    tok = rng.integers(0, V, N).astype(np.int32)
    tok[:8] = tok[8:16]  # repeated tokens: the backward must accumulate
    idx = hv(tok.reshape(N, 1), "i32").expand([N, D])   # the frontend's stride-0 index
    got = B.gather(H.upload(ctx, wte), H.upload(ctx, idx), 0)
    H.assert_same("f32", H.download(got), oracle.gather(wte, idx, 0).numpy(), what="embedding gather")
    g = hv(rng.standard_normal((N, D)).astype(np.float32))
    zeros = hv(np.zeros((V, D), np.float32))
    got = B.scatter(H.upload(ctx, zeros), H.upload(ctx, idx), H.upload(ctx, g), 0, mode="add")
    want = oracle.scatter(zeros, idx, g, 0, "add").numpy()
    close(H.download(got), want, 1e-6, "embedding backward (scatter add)")


def test_gpt2_layernorm_and_softmax_pieces(ctx, oracle):
    rng = np.random.default_rng(32)
    x = hv(rng.standard_normal((N, D)).astype(np.float32))
    tx = H.upload(ctx, x)
    # mean over the feature axis, centred, variance: reduce + column broadcast
    s = B.reduce(tx, "sum", [1])
    ws = oracle.reduce("sum", x, [1])
    close(H.download(s), ws.numpy(), 1e-5, "layernorm sum")
    mean_h = hv(ws.numpy() / D)
    mean_b = mean_h.reshape_contig([N, 1]).expand([N, D])
    cen = B.sub(tx, H.upload(ctx, mean_b))
    H.assert_same("f32", H.download(cen), oracle.binary("sub", x, mean_b).numpy(), ulp=0, what="centre")
    # attention scores [BT, HEADS, T, T]: causal mask by where, softmax over the last axis
    sc = hv(rng.standard_normal((BT, HEADS, T, T)).astype(np.float32) * 3)
    mask = hv(np.tril(np.ones((T, T), np.uint8)).reshape(1, 1, T, T), "bool").expand([BT, HEADS, T, T])
    neg = hv(np.full((1, 1, 1, 1), -1e9, np.float32)).expand([BT, HEADS, T, T])
    tm = B.where(H.upload(ctx, mask), H.upload(ctx, sc), H.upload(ctx, neg))
    wm = oracle.where(mask, sc, neg)
    H.assert_same("f32", H.download(tm), wm.numpy(), ulp=0, what="causal mask")
    mx = B.reduce(tm, "max", [3])
    wmx = oracle.reduce("max", wm, [3])
    H.assert_same("f32", H.download(mx), wmx.numpy(), ulp=0, what="row max")
    mxb = wmx.reshape_contig([BT, HEADS, T, 1]).expand([BT, HEADS, T, T])
    e = B.exp(B.sub(tm, H.upload(ctx, mxb)))
    we = oracle.unary("exp", oracle.binary("sub", wm, mxb))
    H.assert_same("f32", H.download(e), we.numpy(), ulp=2, what="exp(x - max)")
    den = B.reduce(e, "sum", [3])
    close(H.download(den), oracle.reduce("sum", we, [3]).numpy(), 1e-5, "softmax denominator")
    am = B.argmax(tm, 3)
    H.assert_same("i32", H.download(am), oracle.argreduce("argmax", wm, 3).numpy(), what="greedy argmax")


def test_gpt2_attention_and_mlp_products(ctx, oracle):
    rng = np.random.default_rng(33)
    q = hv(rng.standard_normal((BT, HEADS, T, HD)).astype(np.float32))
    k = hv(rng.standard_normal((BT, HEADS, T, HD)).astype(np.float32))
    kt = k.permute([0, 1, 3, 2])                              # the transposed VIEW Rune hands over
    got = B.matmul(H.upload(ctx, q), H.upload(ctx, kt))
    close(H.download(got), oracle.matmul(q, kt).numpy(), 1e-5 * HD ** 0.5, "q k^T")
    p = hv(rng.random((BT, HEADS, T, T)).astype(np.float32))
    got = B.matmul(H.upload(ctx, p), H.upload(ctx, q))
    close(H.download(got), oracle.matmul(p, q).numpy(), 1e-5 * T ** 0.5, "p v")
    x = hv(rng.standard_normal((N, D)).astype(np.float32))
    for name, (kk, nn) in {"qkv": (D, 3 * D), "fc": (D, INNER)}.items():
        w = hv(rng.standard_normal((kk, nn)).astype(np.float32) * 0.02)
        got = B.matmul(H.upload(ctx, x), H.upload(ctx, w))
        close(H.download(got), oracle.matmul(x, w).numpy(), 1e-5 * kk ** 0.5, name)
    h = hv(rng.standard_normal((N, INNER)).astype(np.float32))
    w = hv(rng.standard_normal((INNER, D)).astype(np.float32) * 0.02)
    close(H.download(B.matmul(H.upload(ctx, h), H.upload(ctx, w))), oracle.matmul(h, w).numpy(), 1e-5 * INNER ** 0.5, "proj")
    # gelu (tanh form) is x * 0.5 * (1 + tanh(...)): the transcendental is the piece with a ulp bound
    H.assert_same("f32", H.download(B.tanh(H.upload(ctx, h))), oracle.unary("tanh", h).numpy(), ulp=2, what="tanh")


def test_gpt2_tied_logits_product_and_cross_entropy_pieces(ctx, oracle):
    """logits = h @ wte^T with wte^T a transposed view of the [50257, 768] table: 39 GFLOP, so the
    default f32 mode runs it as 3xTF32 on the tensor cores; the reference binary is the checker."""
    rng = np.random.default_rng(34)
    h = hv(rng.standard_normal((N, D)).astype(np.float32))
    wte = hv(rng.standard_normal((V, D)).astype(np.float32) * 0.02)
    wt = wte.permute([1, 0])
    want = oracle.matmul(h, wt)
    got = B.matmul(H.upload(ctx, h), H.upload(ctx, wt))
    close(H.download(got), want.numpy(), 1e-5 * D ** 0.5, "logits")
    tl = H.upload(ctx, want)
    mx = B.reduce(tl, "max", [1])
    H.assert_same("f32", H.download(mx), oracle.reduce("max", want, [1]).numpy(), ulp=0, what="logit max")
    tgt = hv(rng.integers(0, V, (N, 1)).astype(np.int32), "i32")
    picked = B.gather(tl, H.upload(ctx, tgt), 1)
    H.assert_same("f32", H.download(picked), oracle.gather(want, tgt, 1).numpy(), what="target logit")
    lse_in = B.exp(B.sub(tl, B.expand(B.reshape(mx, [N, 1]), [N, V])))
    wmx = oracle.reduce("max", want, [1]).reshape_contig([N, 1]).expand([N, V])
    wexp = oracle.unary("exp", oracle.binary("sub", want, wmx))
    close(H.download(B.reduce(lse_in, "sum", [1])), oracle.reduce("sum", wexp, [1]).numpy(), 1e-5, "sum exp over the vocabulary")


def test_mlp_4096_layer_forward_backward_ops(ctx, oracle):
    """configs[3] at its width (4096) with a batch the oracle affords (256): forward h = relu(x W + b),
    backward relu mask, dW = x^T g and dx = g W^T through transposed views, db = column sum."""
    rng = np.random.default_rng(35)
    Bn, W = 256, 4096
    x = hv(rng.standard_normal((Bn, W)).astype(np.float32))
    w = hv(rng.standard_normal((W, W)).astype(np.float32) / 64)
    b = hv(rng.standard_normal((1, W)).astype(np.float32) * 0.01).expand([Bn, W])
    tx, tw = H.upload(ctx, x), H.upload(ctx, w)
    pre_w = oracle.binary("add", oracle.matmul(x, w), b)
    pre = B.add(B.matmul(tx, tw), H.upload(ctx, b))
    close(H.download(pre), pre_w.numpy(), 1e-5 * W ** 0.5, "x W + b")
    zero = hv(np.zeros((1, 1), np.float32)).expand([Bn, W])
    tp = H.upload(ctx, pre_w)
    H.assert_same("f32", H.download(B.max(tp, H.upload(ctx, zero))), oracle.binary("max", pre_w, zero).numpy(), ulp=0, what="relu")
    mask = B.cast(B.cmplt(H.upload(ctx, zero), tp), "f32")
    H.assert_same("f32", H.download(mask), oracle.cast(oracle.compare("cmplt", zero, pre_w), "f32").numpy(), ulp=0, what="relu'")
    g = hv(rng.standard_normal((Bn, W)).astype(np.float32))
    tg = H.upload(ctx, g)
    close(H.download(B.matmul(B.permute(tx, [1, 0]), tg)), oracle.matmul(x.permute([1, 0]), g).numpy(), 1e-5 * Bn ** 0.5, "dW")
    close(H.download(B.matmul(tg, B.permute(tw, [1, 0]))), oracle.matmul(g, w.permute([1, 0])).numpy(), 1e-5 * W ** 0.5, "dx")
    close(H.download(B.reduce(tg, "sum", [0])), oracle.reduce("sum", g, [0]).numpy(), 1e-5, "db")
