"""GPU parity of matmul against the oracle through the C ABI: every compute dtype,
transposed lhs / rhs / both, offsets, batched and batch-broadcast operands
(backend_contract.ml:1932-2023, backend_c/test/matmul_test.ml:68-140 fills).

Tolerances: integers bit-exact (accumulation is modular); f64 1e-11*(k+1) and f32
1e-5*(k+1) relative to the row/column magnitude as the contract states; bf16/f16
on the tcgen05 path ELEMENTWISE within one storage ulp of the reference's f32-accumulated,
once-rounded element, plus the reordering slack of an f32 sum of K products where the sum
cancels (assert_gemm_16bit) -- tighter than north_star's 1e-3 relative on every element that
does not cancel, and never relative to the largest element of C.
"""
import numpy as np
import pytest

import raven_b200.backend as B
from raven_b200 import Failure, InvalidArgument
from tests import harness as H

pytestmark = pytest.mark.gpu


def _fill(dtype, m, k, n):
    """The reference's deterministic fills: fa i p = sin((i*k+p)*13 mod 1009), fb p j =
    cos((p*n+j)*7 mod 1013) (matmul_test.ml:68-140)."""
    i, p = np.meshgrid(np.arange(m), np.arange(k), indexing="ij")
    fa = np.sin(((i * k + p) * 13) % 1009)
    p2, j = np.meshgrid(np.arange(k), np.arange(n), indexing="ij")
    fb = np.cos(((p2 * n + j) * 7) % 1013)
    if dtype in H.INTS:
        fa, fb = np.round(fa * 50), np.round(fb * 50)
        if dtype in H.UINTS:
            fa, fb = np.abs(fa), np.abs(fb)
        return fa.astype(np.int64).astype(H.np_storage(dtype)), fb.astype(np.int64).astype(H.np_storage(dtype))
    if dtype in H.COMPLEX:
        return ((fa + 1j * fb[:m, :k] if fb.shape == fa.shape else fa + 0.5j * fa[::-1]).astype(H.np_storage(dtype)),
                (fb - 0.25j * fb[::-1]).astype(H.np_storage(dtype)))
    return H.to_storage(dtype, fa), H.to_storage(dtype, fb)


def _tol(dtype, k):
    if dtype == "f64" or dtype == "c64":
        return 1e-11 * (k + 1)
    if dtype == "f32" or dtype == "c32":
        return 1e-5 * (k + 1) ** 0.5
    return {"f8e4m3": 0.13, "f8e5m2": 0.26}[dtype]


storage_ulp, assert_gemm_16bit = H.storage_ulp, H.assert_gemm_16bit


def _absprod(dtype, a, b):
    fa = np.abs(H.storage_to_float(dtype, a.numpy()))
    fb = np.abs(H.storage_to_float(dtype, b.numpy()))
    return np.matmul(fa, fb)


def _check(ctx, oracle, dtype, a, b, what, k):
    want = oracle.matmul(a, b).numpy()
    got = H.download(B.matmul(H.upload(ctx, a), H.upload(ctx, b)))
    if dtype in H.INTS:
        H.assert_same(dtype, got, want, what=what)
    elif dtype in ("bf16", "f16"):
        assert_gemm_16bit(dtype, got, want, _absprod(dtype, a, b), k, what)
    else:
        scale = float(np.max(np.abs(H.storage_to_float(dtype, want) if dtype not in H.COMPLEX else want))) or 1.0
        H.assert_close(dtype, got, want, rel=_tol(dtype, k), abs_=_tol(dtype, k) * scale, what=what)


ALL = list(H.FLOATS) + list(H.INTS) + list(H.COMPLEX)


@pytest.mark.parametrize("dtype", ALL)
def test_matmul_layouts(ctx, oracle, dtype):
    for (m, k, n) in [(5, 7, 3), (1, 1, 1), (64, 64, 64), (70, 33, 129)]:
        A, Bm = _fill(dtype, m, k, n)
        a = H.HostView(A.reshape(-1).copy(), dtype, [m, k])
        b = H.HostView(Bm.reshape(-1).copy(), dtype, [k, n])
        at = H.HostView(np.ascontiguousarray(A.T).reshape(-1), dtype, [k, m]).permute([1, 0])
        bt = H.HostView(np.ascontiguousarray(Bm.T).reshape(-1), dtype, [n, k]).permute([1, 0])
        for name, (x, y) in {"nn": (a, b), "tn": (at, b), "nt": (a, bt), "tt": (at, bt)}.items():
            _check(ctx, oracle, dtype, x, y, f"matmul/{dtype}/{m}x{k}x{n}/{name}", k)


@pytest.mark.parametrize("dtype", ["f32", "f64", "i32", "bf16", "f16", "c32"])
def test_matmul_batched(ctx, oracle, dtype):
    m, k, n = 9, 17, 6
    A, Bm = _fill(dtype, 4 * m, k, n)
    a = H.HostView(A.reshape(-1).copy(), dtype, [2, 2, m, k])
    b = H.HostView(Bm.reshape(-1).copy(), dtype, [k, n])
    _check(ctx, oracle, dtype, a, b, f"batched-rhs2d/{dtype}", k)
    A2, B2 = _fill(dtype, m, k, 6 * n)
    b3 = H.HostView(np.ascontiguousarray(B2.reshape(k, 6, n).transpose(1, 0, 2)).reshape(-1), dtype, [3, 2, k, n])
    a1 = H.HostView(A2.reshape(-1).copy(), dtype, [1, 1, m, k]).expand([3, 1, m, k])
    _check(ctx, oracle, dtype, a1, b3, f"batch-broadcast/{dtype}", k)
    # offset slice of A
    big = H.HostView(A.reshape(-1).copy(), dtype, [4 * m, k]).shrink([(3, 3 + m), (2, k)])
    bs = H.HostView(Bm.reshape(-1).copy(), dtype, [k, n]).shrink([(2, k), (0, n)])
    _check(ctx, oracle, dtype, big, bs, f"offset/{dtype}", k - 2)


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16", "f64", "i32"])
@pytest.mark.parametrize("shape", [(256, 256, 256), (512, 384, 640), (1024, 1024, 1024), (300, 4100, 200)])
def test_matmul_large(ctx, oracle, dtype, shape):
    m, k, n = shape
    if dtype in ("f64", "i32") and m * k * n > 512 * 384 * 640:
        pytest.skip("oracle too slow at this size for this dtype")
    rng = np.random.default_rng(0)
    if dtype in H.INTS:
        A = rng.integers(-100, 100, (m, k)).astype(np.int32)
        Bm = rng.integers(-100, 100, (k, n)).astype(np.int32)
    else:
        A = H.to_storage(dtype, rng.standard_normal((m, k)) / np.sqrt(k))
        Bm = H.to_storage(dtype, rng.standard_normal((k, n)))
    a = H.HostView(A.reshape(-1).copy(), dtype, [m, k])
    b = H.HostView(Bm.reshape(-1).copy(), dtype, [k, n])
    at = H.HostView(np.ascontiguousarray(A.T).reshape(-1), dtype, [k, m]).permute([1, 0])
    bt = H.HostView(np.ascontiguousarray(Bm.T).reshape(-1), dtype, [n, k]).permute([1, 0])
    for name, (x, y) in {"nn": (a, b), "tn": (at, b), "nt": (a, bt)}.items():
        _check(ctx, oracle, dtype, x, y, f"matmul/{dtype}/{shape}/{name}", k)


def test_matmul_errors(ctx):
    from raven_b200 import dtype as D
    a = B.buffer(ctx, D.float32, [3, 4])
    b = B.buffer(ctx, D.float32, [5, 6])
    with pytest.raises(InvalidArgument, match="matmul: shape mismatch"):
        B.matmul(a, b)
    x = B.buffer(ctx, D.bool_, [2, 2])
    with pytest.raises(Failure, match="matmul: dtype not supported for this operation"):
        B.matmul(x, x)
    # k == 0 is a real zero fill
    z = B.matmul(B.buffer(ctx, D.float32, [3, 0]), B.buffer(ctx, D.float32, [0, 2]))
    assert (H.download(z) == 0).all()


@pytest.fixture
def pair_mode(request):
    """Force the tcgen05 GEMM's tile mode: "1" = 2-CTA pair (cta_group::2, 256x256 tiles),
    "0" = single CTA (128x256). The env override is read at every launch."""
    import os
    old = os.environ.get("NX_CUDA_MM_PAIR")
    os.environ["NX_CUDA_MM_PAIR"] = request.param
    yield request.param
    if old is None:
        del os.environ["NX_CUDA_MM_PAIR"]
    else:
        os.environ["NX_CUDA_MM_PAIR"] = old


TC_SHAPES = [(256, 256, 256), (512, 384, 640), (300, 4100, 200), (128, 64, 256), (1024, 1024, 1024),
             (2048, 512, 1536), (257, 72, 264)]


@pytest.mark.parametrize("pair_mode", ["0", "1"], indirect=True)
@pytest.mark.parametrize("dtype", ["bf16", "f16", "tf32"])
def test_matmul_tc_tile_modes(ctx, oracle, dtype, pair_mode):
    """Both tile modes of the tensor-core GEMM against the oracle: M/N/K tails (a pair whose
    second CTA is entirely past M), all four operand majors, batched + batch-broadcast.
    tf32 = f32 operands in the opt-in tf32 mode, 1e-3 relative (north_star)."""
    st = "f32" if dtype == "tf32" else dtype
    if dtype == "tf32":
        ctx.set_matmul_mode("tf32")
    try:
        rng = np.random.default_rng(1)
        for (m, k, n) in TC_SHAPES:
            A = H.to_storage(st, rng.standard_normal((m, k)) / np.sqrt(k))
            Bm = H.to_storage(st, rng.standard_normal((k, n)))
            a = H.HostView(A.reshape(-1).copy(), st, [m, k])
            b = H.HostView(Bm.reshape(-1).copy(), st, [k, n])
            at = H.HostView(np.ascontiguousarray(A.T).reshape(-1), st, [k, m]).permute([1, 0])
            bt = H.HostView(np.ascontiguousarray(Bm.T).reshape(-1), st, [n, k]).permute([1, 0])
            for name, (x, y) in {"nn": (a, b), "tn": (at, b), "nt": (a, bt), "tt": (at, bt)}.items():
                want = oracle.matmul(x, y).numpy()
                got = H.download(B.matmul(H.upload(ctx, x), H.upload(ctx, y)))
                what = f"{dtype}/pair={pair_mode}/{(m, k, n)}/{name}"
                if dtype in ("bf16", "f16"):
                    assert_gemm_16bit(dtype, got, want, _absprod(dtype, x, y), k, what)
                    continue
                wf, gf = H.storage_to_float(st, want), H.storage_to_float(st, got)
                scale = float(np.max(np.abs(wf))) or 1.0
                tol = 1e-3 if dtype == "tf32" else _tol(dtype, k)
                err = float(np.max(np.abs(gf - wf)))
                assert err <= tol * scale, f"{what}: err {err} scale {scale}"
        # batched: [3, m, k] x [k, n] (broadcast rhs) and [3, m, k] x [3, k, n]
        m, k, n = 384, 96, 320
        A = H.to_storage(st, rng.standard_normal((3, m, k)) / np.sqrt(k))
        Bm = H.to_storage(st, rng.standard_normal((3, k, n)))
        a = H.HostView(A.reshape(-1).copy(), st, [3, m, k])
        for b in (H.HostView(Bm[0].reshape(-1).copy(), st, [k, n]), H.HostView(Bm.reshape(-1).copy(), st, [3, k, n])):
            want_s = oracle.matmul(a, b).numpy()
            got_s = H.download(B.matmul(H.upload(ctx, a), H.upload(ctx, b)))
            if dtype in ("bf16", "f16"):
                assert_gemm_16bit(dtype, got_s, want_s, _absprod(dtype, a, b), k, f"batched/{dtype}")
                continue
            want, got = H.storage_to_float(st, want_s), H.storage_to_float(st, got_s)
            tol = 1e-3 if dtype == "tf32" else _tol(dtype, k)
            assert float(np.max(np.abs(got - want))) <= tol * (float(np.max(np.abs(want))) or 1.0), f"batched/{dtype}"
    finally:
        ctx.set_matmul_mode("f32")


def test_matmul_f32_large_runs_as_3xtf32_and_keeps_f32_accuracy(ctx, oracle):
    """f32 products of >= 0.25 GFLOP take the tensor-core 3xTF32 path (nxc_matmul_x3.cu) by default.
    Against the reference binary's f32 result: 2e-5 of max (|A||B|) -- far inside the reference's own
    f32 tolerance (1e-3 rel + 1e-3 abs, backend_c/test/matmul_test.ml:831) and 50x tighter than plain
    tf32 gets; mode "ieee" pins the CUDA-core kernel and agrees to 2e-6. Transposed views, a K that
    is not a multiple of the k-block, batch broadcast, and an inf operand (no NaN from inf - inf)."""
    rng = np.random.default_rng(11)
    m, k, n = 512, 2056, 1024
    A = rng.standard_normal((m, k)).astype(np.float32)
    Bm = rng.standard_normal((k, n)).astype(np.float32)
    a = H.HostView(A.reshape(-1).copy(), "f32", [m, k])
    b = H.HostView(Bm.reshape(-1).copy(), "f32", [k, n])
    at = H.HostView(np.ascontiguousarray(A.T).reshape(-1), "f32", [k, m]).permute([1, 0])
    bt = H.HostView(np.ascontiguousarray(Bm.T).reshape(-1), "f32", [n, k]).permute([1, 0])
    want = oracle.matmul(a, b).numpy().astype(np.float64)
    bound = float((np.abs(A).astype(np.float64) @ np.abs(Bm).astype(np.float64)).max())
    before = ctx.launch_count()
    for name, (x, y) in {"nn": (a, b), "tn": (at, b), "nt": (a, bt), "tt": (at, bt)}.items():
        got = H.download(B.matmul(H.upload(ctx, x), H.upload(ctx, y))).astype(np.float64)
        assert np.abs(got - want).max() <= 2e-5 * bound, name
    assert ctx.launch_count() - before >= 4 * 3  # two split passes + the GEMM per product
    try:
        ctx.set_matmul_mode("ieee")
        got = H.download(B.matmul(H.upload(ctx, a), H.upload(ctx, bt))).astype(np.float64)
        assert np.abs(got - want).max() <= 2e-6 * bound
    finally:
        ctx.set_matmul_mode("f32")
    # batched A [3, m, k] against a broadcast B [k, n]
    A3 = rng.standard_normal((3, m, k)).astype(np.float32)
    a3 = H.HostView(A3.reshape(-1).copy(), "f32", [3, m, k])
    want3 = oracle.matmul(a3, b).numpy().astype(np.float64)
    got3 = H.download(B.matmul(H.upload(ctx, a3), H.upload(ctx, b))).astype(np.float64)
    assert got3.shape == (3, m, n) and np.abs(got3 - want3).max() <= 2e-5 * bound * 1.5
    # an infinite element gives inf in its row, not NaN
    Ai = A.copy()
    Ai[7, 5] = np.inf
    gi = H.download(B.matmul(H.upload(ctx, H.HostView(Ai.reshape(-1), "f32", [m, k])), H.upload(ctx, b)))
    wi = oracle.matmul(H.HostView(Ai.reshape(-1), "f32", [m, k]), b).numpy()
    assert np.array_equal(np.isinf(gi), np.isinf(wi)) and np.array_equal(np.isnan(gi), np.isnan(wi))
    assert np.array_equal(np.sign(gi[7]), np.sign(wi[7]))


def test_matmul_tensor_core_path_packs_operands_tma_cannot_describe(ctx, oracle):
    """A row pitch that is not a multiple of 16 bytes (GPT-2's 50257-wide logits gradient), a base
    that is not 16-byte aligned (a column slice) and a doubly strided view cannot be handed to TMA:
    large products pack such an operand once (strided copy, K-major, aligned pitch) and still run on
    the tensor cores -- same result as the reference binary to the bf16 tolerance, and more than the
    single launch the CUDA-core fallback would be."""
    rng = np.random.default_rng(12)
    m, k, n = 384, 512, 1001
    A = H.to_storage("bf16", rng.standard_normal((m, k)) / 8)
    Bw = H.to_storage("bf16", rng.standard_normal((k, n + 3)))
    a = H.HostView(A.reshape(-1).copy(), "bf16", [m, k])
    b_full = H.HostView(Bw.reshape(-1).copy(), "bf16", [k, n + 3])
    cases = {
        "odd pitch": (a, H.HostView(np.ascontiguousarray(Bw[:, :n]).reshape(-1), "bf16", [k, n])),
        "column slice (misaligned base, odd pitch)": (a, b_full.shrink([(0, k), (1, n + 1)])),
        "odd-pitch lhs through a transposed view": (
            H.HostView(np.ascontiguousarray(A.T[:, :m - 1]).reshape(-1), "bf16", [k, m - 1]).permute([1, 0]),
            H.HostView(np.ascontiguousarray(Bw[:, :n]).reshape(-1), "bf16", [k, n])),
    }
    for name, (x, y) in cases.items():
        want = oracle.matmul(x, y).numpy()
        before = ctx.launch_count()
        got = H.download(B.matmul(H.upload(ctx, x), H.upload(ctx, y)))
        assert ctx.launch_count() - before >= 2, name
        assert_gemm_16bit("bf16", got, want, _absprod("bf16", x, y), k, name)
    # batched lhs whose batch stride is not a 16-byte multiple, against a broadcast rhs
    A3 = H.to_storage("bf16", rng.standard_normal((3, 128, k + 1)) / 8)
    a3 = H.HostView(A3.reshape(-1).copy(), "bf16", [3, 128, k + 1]).shrink([(0, 3), (0, 128), (0, k)])
    y = cases["odd pitch"][1]
    want = oracle.matmul(a3, y).numpy()
    got = H.download(B.matmul(H.upload(ctx, a3), H.upload(ctx, y)))
    assert_gemm_16bit("bf16", got, want, _absprod("bf16", a3, y), k, "batched odd pitch")


def test_matmul_f32_attention_shaped_batch_runs_as_3xtf32(ctx, oracle):
    """q k^T at GPT-2 shapes with a full-length sequence: [2, 12, 1024, 64] x a transposed VIEW of
    [2, 12, 1024, 64] -- 3.2 GFLOP, K = 64 (two k-blocks per section), two batch dims that collapse to
    one stride. Against the reference binary's f32 product."""
    rng = np.random.default_rng(13)
    Bt, Hh, T, Dh = 2, 12, 1024, 64
    q = H.HostView.from_array(rng.standard_normal((Bt, Hh, T, Dh)).astype(np.float32), "f32")
    k = H.HostView.from_array(rng.standard_normal((Bt, Hh, T, Dh)).astype(np.float32), "f32")
    kt = k.permute([0, 1, 3, 2])
    want = oracle.matmul(q, kt).numpy().astype(np.float64)
    before = ctx.launch_count()
    got = H.download(B.matmul(H.upload(ctx, q), H.upload(ctx, kt))).astype(np.float64)
    assert ctx.launch_count() - before == 3
    bound = float(np.abs(want).max()) * 4
    assert got.shape == (Bt, Hh, T, T) and np.abs(got - want).max() <= 2e-5 * bound


def test_matmul_m_major_lhs_with_wide_n_is_packed_and_bit_identical(ctx):
    """dW = x^T g shapes: an M-major (transposed-view) left operand with N >= 4096 is transposed once
    into a K-major buffer before the tensor-core kernel (nxc_matmul.cu). Property check at a size the
    oracle cannot afford: the result equals, bit for bit, the product of the materialised transpose,
    and differs in launch count from the in-place route (NX_CUDA_MM_NO_APACK is not set here)."""
    rng = np.random.default_rng(14)
    m, k, n = 1024, 4096, 4096
    xt = B.cast(B.reshape(B.from_host(ctx, rng.standard_normal(k * m).astype(np.float32)), [k, m]), "bf16")   # x: [k, m]
    g = B.cast(B.reshape(B.from_host(ctx, rng.standard_normal(k * n).astype(np.float32)), [k, n]), "bf16")
    before = ctx.launch_count()
    via_view = H.download(B.matmul(B.permute(xt, [1, 0]), g))
    n_view = ctx.launch_count() - before
    a = B.contiguous(B.permute(xt, [1, 0]))
    before = ctx.launch_count()
    via_copy = H.download(B.matmul(a, g))
    assert ctx.launch_count() - before == 1 and n_view == 2
    assert via_view.shape == (m, n) and np.array_equal(via_view, via_copy)
    # and a spot check of values against float64 on a few rows
    xf = H.bf16_bits_to_f32(H.download(xt)).astype(np.float64)
    gf = H.bf16_bits_to_f32(H.download(g)).astype(np.float64)
    want = xf[:, :4].T @ gf
    got = H.bf16_bits_to_f32(via_view[:4]).astype(np.float64)
    bound = storage_ulp("bf16", want) + 2.0 * k * 2.0 ** -24 * (np.abs(xf[:, :4]).T @ np.abs(gf))
    assert (np.abs(got - want) <= bound).all(), float(np.max(np.abs(got - want) / bound))


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_matmul_broadcast_weight_batches_fold_into_rows(ctx, oracle, dtype):
    """x [B, T, C] @ W [C, C'] (every Linear.apply of the GPT-2 step): the batches follow each other at
    the row pitch, so the engine runs ONE product with B*T rows. Same answer as the reference's
    batched loop; a batch that does NOT compose with the rows (a sliced batch) keeps the batched route."""
    rng = np.random.default_rng(21)
    Bt, T, C, C2 = 4, 64, 768, 768
    X = H.to_storage(dtype, rng.standard_normal((Bt, T, C)) / 4)
    W = H.to_storage(dtype, rng.standard_normal((C, C2)) / 16)
    x = H.HostView(X.reshape(-1).copy(), dtype, [Bt, T, C])
    w = H.HostView(W.reshape(-1).copy(), dtype, [C, C2])
    x_sliced = H.HostView(np.tile(X.reshape(-1), 2), dtype, [2 * Bt, T, C]).shrink([(0, 2 * Bt), (0, T), (0, C)])
    x_sliced = H.HostView(x_sliced.storage, dtype, [Bt, T, C], [2 * T * C, C, 1], 0)     # every other batch
    for name, xv in (("dense", x), ("every other batch", x_sliced)):
        want = oracle.matmul(xv, w).numpy()
        got = H.download(B.matmul(H.upload(ctx, xv), H.upload(ctx, w)))
        assert got.shape == (Bt, T, C2)
        if dtype == "bf16":
            assert_gemm_16bit(dtype, got, want, _absprod(dtype, xv, w), C, name)
        else:
            bound = float((np.abs(X.reshape(-1, C)).astype(np.float64) @ np.abs(W).astype(np.float64)).max())
            assert np.abs(got.astype(np.float64) - want.astype(np.float64)).max() <= 2e-5 * bound, name


def test_matmul_f32_short_wide_k_runs_its_three_sections_as_batches(ctx, oracle, monkeypatch):
    """A 3xTF32 product with few output tiles and K >= 256 (a 256 x 768 x 768 linear of the GPT-2
    step at the reference's batch) runs its lo*hi, hi*lo and hi*hi sections as three batches and
    adds the partials with one fold (nxc_matmul_x3.cu): same accuracy class against the reference
    binary, the same result as the single-chain form to rounding, strided output rows, a K that is
    not a multiple of 32, and an infinite operand."""
    rng = np.random.default_rng(12)
    for m, k, n in ((256, 768, 768), (64, 4104, 512), (300, 2048, 256)):
        A = rng.standard_normal((m, k)).astype(np.float32)
        Bm = rng.standard_normal((k, n)).astype(np.float32)
        a = H.HostView(A.reshape(-1).copy(), "f32", [m, k])
        b = H.HostView(Bm.reshape(-1).copy(), "f32", [k, n])
        bt = H.HostView(np.ascontiguousarray(Bm.T).reshape(-1), "f32", [n, k]).permute([1, 0])
        want = oracle.matmul(a, b).numpy().astype(np.float64)
        bound = float((np.abs(A).astype(np.float64) @ np.abs(Bm).astype(np.float64)).max())
        da, db = H.upload(ctx, a), H.upload(ctx, bt)
        before = ctx.launch_count()
        got = H.download(B.matmul(da, db)).astype(np.float64)
        split_launches = ctx.launch_count() - before
        assert np.abs(got - want).max() <= 2e-5 * bound, (m, k, n)
        monkeypatch.setenv("NX_CUDA_X3_NO_SPLIT", "1")
        before = ctx.launch_count()
        chain = H.download(B.matmul(da, db)).astype(np.float64)
        chain_launches = ctx.launch_count() - before
        monkeypatch.delenv("NX_CUDA_X3_NO_SPLIT")
        assert chain_launches == 3 and split_launches > chain_launches, (m, k, n)   # + the fold of the partials
        assert np.abs(got - chain).max() <= 2e-6 * bound, (m, k, n)
    Ai = A.copy()
    Ai[3, 9] = -np.inf
    gi = H.download(B.matmul(H.upload(ctx, H.HostView(Ai.reshape(-1), "f32", [m, k])), H.upload(ctx, b)))
    wi = oracle.matmul(H.HostView(Ai.reshape(-1), "f32", [m, k]), b).numpy()
    assert np.array_equal(np.isinf(gi), np.isinf(wi)) and np.array_equal(np.isnan(gi), np.isnan(wi))
    assert np.array_equal(np.sign(gi[3]), np.sign(wi[3]))
