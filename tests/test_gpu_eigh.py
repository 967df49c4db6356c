"""GPU parity of eigh / eigvalsh (linalg tier 2) through the C ABI:
eigenvalues against the oracle (ascending, always f64), eigenvectors by the defining properties
(unique only up to a phase per column), batched, strided, lower-triangle-only, error classes."""
import numpy as np
import pytest

import raven_b200.backend as B
from raven_b200 import InvalidArgument
from tests import harness as H
from tests.golden.make_golden_eigh import inputs
from tests.test_oracle_eigh import TOL, check_eigh

pytestmark = pytest.mark.gpu


def test_eigh_matches_oracle(ctx, oracle):
    for key, hv in inputs():
        dt = key.split("|")[1]
        want = oracle.eigh(hv, False).numpy()
        w, v = B.eigh(H.upload(ctx, hv))
        w, v = H.download(w), H.download(v)
        assert w.dtype == np.float64 and w.shape == want.shape, key
        assert np.abs(w - want).max() <= 20 * TOL[dt] * max(1.0, np.abs(want).max()), key
        check_eigh(w, v, hv, dt, key, tol_scale=2.0)
        wv = H.download(B.eigvalsh(H.upload(ctx, hv)))
        assert np.abs(wv - want).max() <= 20 * TOL[dt] * max(1.0, np.abs(want).max()), key


def test_eigh_reads_lower_triangle_and_strided_views(ctx, oracle):
    rng = np.random.default_rng(52)
    n = 7
    a = rng.standard_normal((n, n))
    sym = a + a.T
    junk = np.tril(sym) + np.triu(rng.standard_normal((n, n)), 1)   # upper triangle is garbage
    want = np.linalg.eigvalsh(sym)
    w = H.download(B.eigvalsh(H.upload(ctx, H.HostView.from_array(junk, "f64"))))
    assert np.abs(w - want).max() <= 1e-11 * np.abs(want).max()
    hv = H.HostView.from_array(sym, "f64").flip([True, True])
    w = H.download(B.eigvalsh(H.upload(ctx, hv)))
    assert np.abs(w - want).max() <= 1e-11 * np.abs(want).max()


def test_eigh_larger_residual(ctx):
    rng = np.random.default_rng(53)
    n = 96
    a = rng.standard_normal((3, n, n)) + 1j * rng.standard_normal((3, n, n))
    h = a + np.conj(np.swapaxes(a, -1, -2))
    w, v = B.eigh(H.upload(ctx, H.HostView.from_array(h, "c64")))
    w, v = H.download(w), H.download(v)
    assert np.abs(w - np.linalg.eigvalsh(h)).max() <= 1e-10 * np.abs(w).max()
    assert np.abs(h @ v - v * w[:, None, :]).max() <= 1e-9 * np.abs(w).max()
    assert np.abs(np.conj(np.swapaxes(v, -1, -2)) @ v - np.eye(n)).max() <= 1e-10


def test_eigh_errors(ctx):
    up = lambda a, dt: H.upload(ctx, H.HostView.from_array(a, dt))
    with pytest.raises(InvalidArgument, match="eigh: matrix must be square"):
        B.eigh(up(np.ones((2, 3)), "f64"))
    with pytest.raises(InvalidArgument, match="eigvalsh: linalg requires a float or complex dtype"):
        B.eigvalsh(up(np.ones((2, 2), dtype=np.int32), "i32"))


@pytest.mark.parametrize("dt", ["f32", "f64", "c64"])
def test_eigh_cluster_teams(ctx, dt, monkeypatch):
    """Matrices with >= 64 column pairs per step are rotated by a thread-block cluster (nxc_linalg.cu,
    nxc_la3_team): every team width gives the decomposition the one-CTA launch gives (same
    schedule, same arithmetic per pair, so eigenvalues agree to rounding), batched as well."""
    rng = np.random.default_rng(54)
    n = 200
    a = rng.standard_normal((2, n, n)) + (1j * rng.standard_normal((2, n, n)) if dt == "c64" else 0)
    h = a + np.conj(np.swapaxes(a, -1, -2))
    npdt = {"f32": np.float32, "f64": np.float64, "c64": np.complex128}[dt]
    hv = H.HostView.from_array(h.astype(npdt), dt)
    want = np.linalg.eigvalsh(h)
    tol = 2e-5 if dt == "f32" else 1e-11
    for width in ("1", "2", "8", "16", None):
        if width is None:
            monkeypatch.delenv("NX_CUDA_LA_CLUSTER", raising=False)
        else:
            monkeypatch.setenv("NX_CUDA_LA_CLUSTER", width)
        w, v = B.eigh(H.upload(ctx, hv))
        w, v = H.download(w), H.download(v).astype(np.complex128 if dt == "c64" else np.float64)
        assert np.abs(w - want).max() <= tol * np.abs(want).max(), width
        assert np.abs(h @ v - v * w[:, None, :]).max() <= 20 * tol * np.abs(want).max(), width
        assert np.abs(np.conj(np.swapaxes(v, -1, -2)) @ v - np.eye(n)).max() <= 20 * tol, width
