"""GPU parity of linalg tier 1 against the oracle through the C
ABI: cholesky (lower / upper), triangular_solve (upper x conjugate-transpose x unit-diagonal,
matrix and vector right-hand sides), qr (reduced / full, tall / wide / square), batched, every
float / complex dtype the reference accepts (f16 / bf16 compute in f32), strided operands, the
reference's failure classes. Tolerances are relative to the largest output magnitude; larger
single matrices are checked through residuals (A = L L^H, op(A) X = B, A = Q R, Q^H Q = I).
"""
import numpy as np
import pytest

import raven_b200.backend as B
from raven_b200 import Failure, InvalidArgument
from tests import harness as H
from tests.golden.make_golden_linalg import WIDE, _mk, hv_of

pytestmark = pytest.mark.gpu

TOL = {"f32": 4e-4, "f64": 1e-11, "c32": 4e-4, "c64": 1e-11, "bf16": 6e-2, "f16": 8e-3}
DTS = ["f32", "f64", "c32", "c64", "bf16", "f16"]


def _wide(oracle, hv):
    return oracle.cast(hv, WIDE[hv.dtype]).numpy()


def _dev_wide(oracle, t, dt):
    arr = H.download(t)
    return _wide(oracle, H.HostView.from_array(arr, dt) if arr.dtype != np.uint16 else H.HostView(arr.reshape(-1), dt, list(arr.shape)))


def _close(got, want, dt, what, scale=1.0):
    assert got.shape == want.shape, what
    if want.size:
        err = np.abs(got - want).max() / max(1.0, np.abs(want).max())
        assert err <= TOL[dt] * scale, f"{what}: {err:.3e}"


@pytest.mark.parametrize("dt", DTS)
def test_cholesky_and_solve(ctx, oracle, dt):
    rng = np.random.default_rng(41)
    for bshape, n in (((), 1), ((), 4), ((2,), 7), ((2, 3), 5), ((), 40)):
        a = _mk(rng, bshape + (n, n), dt)
        spd = hv_of(a @ np.conj(np.swapaxes(a, -1, -2)) + n * np.eye(n), dt)
        for upper in (False, True):
            want = _wide(oracle, oracle.cholesky(spd, upper))
            got = _dev_wide(oracle, B.cholesky(H.upload(ctx, spd), upper), dt)
            _close(got, want, dt, f"cholesky/{dt}/{bshape}/{n}/upper={upper}")
        tri = {False: hv_of(np.tril(a) / n + np.eye(n), dt), True: hv_of(np.triu(a) / n + np.eye(n), dt)}
        for nrhs in (1, 3):
            b = hv_of(_mk(rng, bshape + (n, nrhs), dt), dt)
            for upper, tr, unit in ((False, False, False), (False, True, False), (True, False, True),
                                    (True, True, False), (False, False, True)):
                want = _wide(oracle, oracle.triangular_solve(tri[upper], b, upper, tr, unit))
                got = _dev_wide(oracle, B.triangular_solve(H.upload(ctx, tri[upper]), H.upload(ctx, b), upper, tr, unit), dt)
                _close(got, want, dt, f"trsm/{dt}/{bshape}/{n}/{nrhs}/{upper}{tr}{unit}", scale=4)
        # vector right-hand side (rank one less than A)
        bv = hv_of(_mk(rng, bshape + (n,), dt), dt)
        want = _wide(oracle, oracle.triangular_solve(tri[False], bv, False, False, False))
        got = _dev_wide(oracle, B.triangular_solve(H.upload(ctx, tri[False]), H.upload(ctx, bv)), dt)
        _close(got, want, dt, f"trsm-vector/{dt}/{bshape}/{n}", scale=4)


@pytest.mark.parametrize("dt", DTS)
def test_qr(ctx, oracle, dt):
    rng = np.random.default_rng(42)
    for bshape, m, n in (((), 4, 4), ((2,), 6, 3), ((), 3, 6), ((2,), 1, 1), ((), 30, 20), ((3,), 8, 8)):
        x = hv_of(_mk(rng, bshape + (m, n), dt), dt)
        for red in (True, False):
            wq, wr = oracle.qr(x, red)
            gq, gr = B.qr(H.upload(ctx, x), red)
            _close(_dev_wide(oracle, gq, dt), _wide(oracle, wq), dt, f"qr.Q/{dt}/{bshape}/{m}x{n}/{red}", scale=4)
            _close(_dev_wide(oracle, gr, dt), _wide(oracle, wr), dt, f"qr.R/{dt}/{bshape}/{m}x{n}/{red}", scale=4)


def test_linalg_strided_operands(ctx, oracle):
    rng = np.random.default_rng(43)
    n = 6
    a = rng.standard_normal((n, n))
    spd = H.HostView.from_array(a @ a.T + n * np.eye(n), "f64")
    for name, v in {"T": spd.permute([1, 0]), "flip": spd.flip([True, True])}.items():
        want = oracle.cholesky(v, False).numpy()
        got = H.download(B.cholesky(H.upload(ctx, v), False))
        _close(got, want, "f64", f"cholesky/strided/{name}")
    x = H.HostView.from_array(rng.standard_normal((5, 8)), "f64").permute([1, 0])
    wq, wr = oracle.qr(x, True)
    gq, gr = B.qr(H.upload(ctx, x), True)
    _close(H.download(gq), wq.numpy(), "f64", "qr/strided/Q", scale=4)
    _close(H.download(gr), wr.numpy(), "f64", "qr/strided/R", scale=4)


def test_linalg_residuals_large(ctx):
    """Sizes beyond what the oracle checks quickly: residual properties in f64."""
    rng = np.random.default_rng(44)
    n = 300
    a = rng.standard_normal((2, n, n))
    spd = a @ np.swapaxes(a, -1, -2) + n * np.eye(n)
    L = H.download(B.cholesky(H.upload(ctx, H.HostView.from_array(spd, "f64"))))
    assert np.abs(L @ np.swapaxes(L, -1, -2) - spd).max() <= 1e-10 * np.abs(spd).max()
    assert np.abs(np.triu(L[0], 1)).max() == 0.0
    b = rng.standard_normal((2, n, 5))
    X = H.download(B.triangular_solve(H.upload(ctx, H.HostView.from_array(L, "f64")), H.upload(ctx, H.HostView.from_array(b, "f64"))))
    assert np.abs(L @ X - b).max() <= 1e-9 * max(1.0, np.abs(b).max())
    m = rng.standard_normal((200, 120))
    q, r = B.qr(H.upload(ctx, H.HostView.from_array(m, "f64")), True)
    q, r = H.download(q), H.download(r)
    assert np.abs(q @ r - m).max() <= 1e-11 * n
    assert np.abs(q.T @ q - np.eye(120)).max() <= 1e-12 * n
    assert np.abs(np.tril(r, -1)).max() == 0.0


def test_linalg_errors(ctx):
    up = lambda a, dt: H.upload(ctx, H.HostView.from_array(a, dt))
    with pytest.raises(Failure, match="cholesky: matrix is not positive definite") as e:
        B.cholesky(up(-np.eye(3), "f64"))
    assert isinstance(e.value, B.LinalgError) and e.value.kind == "Not_positive_definite"
    with pytest.raises(Failure, match="cholesky: matrix is not positive definite"):
        B.cholesky(up(np.full((2, 2), np.nan), "f32"))
    with pytest.raises(InvalidArgument, match="cholesky: matrix must be square"):
        B.cholesky(up(np.ones((2, 3)), "f64"))
    with pytest.raises(InvalidArgument, match="cholesky: linalg requires a float or complex dtype"):
        B.cholesky(up(np.ones((2, 2), dtype=np.int32), "i32"))
    with pytest.raises(Failure, match="triangular_solve: triangular matrix is singular") as e:
        B.triangular_solve(up(np.zeros((3, 3)), "f64"), up(np.ones((3, 2)), "f64"))
    assert e.value.kind == "Singular"
    with pytest.raises(InvalidArgument, match="triangular_solve: operand shapes are incompatible"):
        B.triangular_solve(up(np.eye(3), "f64"), up(np.ones((4, 2)), "f64"))
    # a unit-diagonal solve never looks at the (zero) diagonal
    x = H.download(B.triangular_solve(up(np.zeros((3, 3)), "f64"), up(np.ones((3, 2)), "f64"), unit_diag=True))
    assert np.array_equal(x, np.ones((3, 2)))


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_cholesky_blocked_large(ctx, oracle, dt, monkeypatch):
    """Large real matrices take the panel-blocked path (nxc_linalg.cu, nxc_cholesky_blocked): checked
    against the oracle at 320 (ragged last panel) and 512, lower and upper, a small batch, against
    the one-CTA kernel on the same input, and through the residual at 2048."""
    rng = np.random.default_rng(45)
    npdt = np.float32 if dt == "f32" else np.float64
    for bshape, n in (((), 320), ((2,), 512)):
        a = rng.standard_normal(bshape + (n, n))
        spd = hv_of((a @ np.swapaxes(a, -1, -2) / n + np.eye(n)).astype(npdt), dt)
        for upper in (False, True):
            want = _wide(oracle, oracle.cholesky(spd, upper))
            got = _dev_wide(oracle, B.cholesky(H.upload(ctx, spd), upper), dt)
            _close(got, want, dt, f"cholesky-blocked/{dt}/{bshape}/{n}/upper={upper}")
            other = np.triu(got, 1) if not upper else np.tril(got, -1)
            assert np.abs(other).max() == 0.0
        monkeypatch.setenv("NX_CUDA_CHOLESKY_BLOCKED", "0")
        one_cta = _dev_wide(oracle, B.cholesky(H.upload(ctx, spd), False), dt)
        monkeypatch.delenv("NX_CUDA_CHOLESKY_BLOCKED")
        _close(_dev_wide(oracle, B.cholesky(H.upload(ctx, spd), False), dt), one_cta, dt, f"blocked vs one-CTA/{dt}/{n}")
    n = 2048
    a = rng.standard_normal((n, n))
    spd = (a @ a.T / n + np.eye(n)).astype(npdt)
    L = H.download(B.cholesky(H.upload(ctx, H.HostView.from_array(spd, dt)))).astype(np.float64)
    assert np.abs(L @ L.T - spd).max() <= (2e-5 if dt == "f32" else 1e-12) * np.abs(spd).max()
    # a matrix that stops being positive definite in a late panel is reported like a small one
    bad = spd.copy()
    bad[1500, 1500] = -1.0
    with pytest.raises(Failure, match="cholesky: matrix is not positive definite"):
        B.cholesky(H.upload(ctx, H.HostView.from_array(bad, dt)))


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_triangular_solve_blocked_large(ctx, oracle, dt):
    """n >= 128 real systems take the panel-blocked substitution (nxc_trsm_blocked): every flag
    combination against the oracle at a ragged size, matrix / single-column right-hand sides,
    batched, and the singular report from a late block."""
    rng = np.random.default_rng(46)
    npdt = np.float32 if dt == "f32" else np.float64
    for bshape, n, nrhs in (((), 200, 70), ((2,), 256, 1), ((), 130, 300)):
        a = rng.standard_normal(bshape + (n, n))
        tri = {False: hv_of((np.tril(a) / n + np.eye(n)).astype(npdt), dt), True: hv_of((np.triu(a) / n + np.eye(n)).astype(npdt), dt)}
        b = hv_of(rng.standard_normal(bshape + (n, nrhs)).astype(npdt), dt)
        for upper in (False, True):
            for tr in (False, True):
                for unit in (False, True):
                    want = _wide(oracle, oracle.triangular_solve(tri[upper], b, upper, tr, unit))
                    got = _dev_wide(oracle, B.triangular_solve(H.upload(ctx, tri[upper]), H.upload(ctx, b), upper, tr, unit), dt)
                    _close(got, want, dt, f"trsm-blocked/{dt}/{bshape}/{n}/{nrhs}/{upper}{tr}{unit}", scale=4)
    n = 300
    sing = np.tril(rng.standard_normal((n, n))) / n + np.eye(n)
    sing[250, 250] = 0.0
    rhs = H.upload(ctx, H.HostView.from_array(np.ones((n, 2), dtype=npdt), dt))
    with pytest.raises(Failure, match="triangular_solve: triangular matrix is singular"):
        B.triangular_solve(H.upload(ctx, H.HostView.from_array(sing.astype(npdt), dt)), rhs)
    # the other triangle is never read
    junk = sing.copy()
    junk[250, 250] = 1.0
    clean = H.download(B.triangular_solve(H.upload(ctx, H.HostView.from_array(junk.astype(npdt), dt)), rhs))
    junk += np.triu(np.full((n, n), np.nan), 1)
    assert np.array_equal(H.download(B.triangular_solve(H.upload(ctx, H.HostView.from_array(junk.astype(npdt), dt)), rhs)), clean)


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_qr_blocked_large(ctx, oracle, dt, monkeypatch):
    """min(m, n) >= 128 real matrices take the compact-WY blocked path (nxc_qr_blocked): against the
    oracle (same reflector sign convention, so Q and R compare directly) for tall / wide / square /
    ragged shapes, reduced and full, batched; against the one-CTA kernel; residuals at 1024."""
    rng = np.random.default_rng(47)
    npdt = np.float32 if dt == "f32" else np.float64
    for bshape, m, n in (((), 200, 200), ((2,), 300, 150), ((), 140, 260), ((), 257, 129)):
        x = hv_of(rng.standard_normal(bshape + (m, n)).astype(npdt), dt)
        for red in (True, False):
            wq, wr = oracle.qr(x, red)
            gq, gr = B.qr(H.upload(ctx, x), red)
            _close(_dev_wide(oracle, gq, dt), _wide(oracle, wq), dt, f"qr-blocked.Q/{dt}/{bshape}/{m}x{n}/{red}", scale=8)
            _close(_dev_wide(oracle, gr, dt), _wide(oracle, wr), dt, f"qr-blocked.R/{dt}/{bshape}/{m}x{n}/{red}", scale=8)
            r = _dev_wide(oracle, gr, dt)
            assert np.abs(np.tril(r.reshape((-1,) + r.shape[-2:])[0], -1)).max() == 0.0
    # both panel kernels (the cluster one is chosen for tall panels of small batches) against each other
    x = hv_of(rng.standard_normal((2, 600, 200)).astype(npdt), dt)
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("NX_CUDA_QR_CLUSTER", mode)
        gq, gr = B.qr(H.upload(ctx, x), True)
        outs[mode] = (_dev_wide(oracle, gq, dt), _dev_wide(oracle, gr, dt))
    monkeypatch.delenv("NX_CUDA_QR_CLUSTER")
    _close(outs["1"][0], outs["0"][0], dt, f"qr cluster vs one-CTA panel Q/{dt}", scale=4)
    _close(outs["1"][1], outs["0"][1], dt, f"qr cluster vs one-CTA panel R/{dt}", scale=4)
    wq, wr = oracle.qr(x, True)
    _close(outs["1"][1], _wide(oracle, wr), dt, f"qr cluster panel R vs oracle/{dt}", scale=8)
    # a column that is already reduced (tau = 0) inside a panel, and a rank-deficient block
    z = rng.standard_normal((200, 160)).astype(npdt)
    z[41:, 40] = 0.0
    z[:, 100:110] = 0.0
    x = hv_of(z, dt)
    wq, wr = oracle.qr(x, True)
    gq, gr = B.qr(H.upload(ctx, x), True)
    q, r = _dev_wide(oracle, gq, dt), _dev_wide(oracle, gr, dt)
    tol = 2e-4 if dt == "f32" else 1e-11
    assert np.abs(q @ r - z).max() <= tol * 10 and np.abs(q.T @ q - np.eye(160)).max() <= tol * 10
    _close(r, _wide(oracle, wr), dt, f"qr-blocked.R/reduced columns/{dt}", scale=8)
    m = 1024
    a = rng.standard_normal((m, m)).astype(npdt)
    q, r = B.qr(H.upload(ctx, H.HostView.from_array(a, dt)), True)
    q, r = H.download(q).astype(np.float64), H.download(r).astype(np.float64)
    assert np.abs(q @ r - a).max() <= tol * 40
    assert np.abs(q.T @ q - np.eye(m)).max() <= tol * 40
