"""GPU parity of the fft family against the oracle through the C
ABI: fft / ifft / rfft / irfft, unnormalised, any length (powers of two in shared memory, longer
ones by the four-step route, every other length by Bluestein's chirp-z), one and several axes,
strided inputs, explicit irfft sizes. Both sides compute in double, so the tolerance is a small
multiple of the OUTPUT type's epsilon relative to the largest output magnitude: 1e-5 for c32 /
f32 results (the north star's reduction bound), 1e-11 for c64 / f64.
"""
import numpy as np
import pytest

import raven_b200.backend as B
from raven_b200 import Failure, InvalidArgument
from tests import harness as H

pytestmark = pytest.mark.gpu

TOL = {"c32": 1e-5, "c64": 1e-11, "f32": 1e-5, "f64": 1e-11}


def _mk(rng, shape, dt):
    if dt in ("c32", "c64"):
        a = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    else:
        a = rng.standard_normal(shape)
    return H.HostView.from_array(a, dt)


def _close(got, want, dt, what):
    assert got.shape == want.shape, what
    if want.size == 0:
        return
    scale = max(1.0, float(np.abs(want).max()))
    err = float(np.abs(got.astype(np.complex128) - want.astype(np.complex128)).max()) / scale
    assert err <= TOL[dt], f"{what}: relative error {err:.3e} > {TOL[dt]:.1e}"


CASES = [((1,), [0]), ((2,), [0]), ((8,), [0]), ((12,), [0]), ((17,), [0]), ((100,), [0]), ((256,), [0]),
         ((4, 6), [0, 1]), ((3, 5, 8), [2]), ((3, 5, 8), [0, 2]), ((5, 7), [1, 0]), ((6, 64), [1]),
         ((33, 4), [0])]


@pytest.mark.parametrize("dt", ["c32", "c64"])
def test_fft_ifft_small(ctx, oracle, dt):
    rng = np.random.default_rng(11)
    for shape, axes in CASES:
        x = _mk(rng, shape, dt)
        for inv in (False, True):
            want = oracle.fft(x, axes, inv).numpy()
            got = H.download((B.ifft if inv else B.fft)(H.upload(ctx, x), axes))
            _close(got, want, dt, f"fft/{dt}/{shape}/{axes}/inv={inv}")


@pytest.mark.parametrize("dt", ["c32", "c64"])
def test_fft_strided_views(ctx, oracle, dt):
    rng = np.random.default_rng(12)
    base = _mk(rng, (6, 10), dt)
    for name, v in {"T": base.permute([1, 0]), "slice": base.shrink([(1, 5), (2, 10)]),
                    "flip": base.flip([True, False])}.items():
        for axes in ([0], [1], [0, 1]):
            want = oracle.fft(v, axes, False).numpy()
            got = H.download(B.fft(H.upload(ctx, v), axes))
            _close(got, want, dt, f"fft/{dt}/{name}/{axes}")


def test_fft_long_lines_match_reference(ctx, oracle):
    """Lengths beyond the shared-memory core (four-step route) and a long Bluestein line; the
    oracle here must be the reference binary (the O(n^2) restatement would take minutes)."""
    if oracle.__name__.endswith("nxo"):
        pytest.skip("reference binary not available")
    rng = np.random.default_rng(13)
    for shape, axes in (((2, 8192), [1]), ((16384,), [0]), ((3, 5000), [1]), ((4099,), [0]), ((4096, 3), [0])):
        x = _mk(rng, shape, "c64")
        want = oracle.fft(x, axes, False).numpy()
        got = H.download(B.fft(H.upload(ctx, x), axes))
        _close(got, want, "c64", f"fft/long/{shape}")
        back = H.download(B.ifft(B.fft(H.upload(ctx, x), axes), axes))
        n = np.prod([shape[a] for a in axes])
        _close(back / n, x.numpy(), "c64", f"ifft(fft)/long/{shape}")


def test_fft_four_step_lines(ctx, oracle):
    """Lines beyond 4096 points go through the lines kernel twice (columns, then twiddled rows with
    a transposed write -- nxc_fft.cu, nxc_fft_pow2): even and odd log2 splits, batched, along a
    strided axis, c32 storage, a long Bluestein length (its power-of-two legs take the same
    route), rfft / irfft with the Hermitian rebuild and a truncated output, and in-place
    multi-axis -- against the reference binary, and numpy's pocketfft for the longest line."""
    if oracle.__name__.endswith("nxo"):
        pytest.skip("reference binary not available")
    rng = np.random.default_rng(16)
    for dt, shape, axes in (("c64", (3, 8192), [1]), ("c64", (2, 32768), [1]), ("c64", (8192, 5), [0]),
                            ("c32", (2, 16384), [1]), ("c64", (2, 9001), [1]), ("c64", (8192, 8), [0, 1])):
        x = _mk(rng, shape, dt)
        want = oracle.fft(x, axes, False).numpy()
        got = H.download(B.fft(H.upload(ctx, x), axes))
        _close(got, want, dt, f"fft/four-step/{dt}/{shape}/{axes}")
        want = oracle.fft(x, axes, True).numpy()
        got = H.download(B.ifft(H.upload(ctx, x), axes))
        _close(got, want, dt, f"ifft/four-step/{dt}/{shape}/{axes}")
    for rdt, cdt, shape in (("f64", "c64", (2, 16384)), ("f32", "c32", (3, 8192))):
        x = _mk(rng, shape, rdt)
        want = oracle.rfft(x, cdt, [1])
        got = H.download(B.rfft(H.upload(ctx, x), cdt, [1]))
        _close(got, want.numpy(), cdt, f"rfft/four-step/{shape}")
        for s in (None, [shape[1] - 100], [shape[1]]):
            w = oracle.irfft(want, rdt, [1], s).numpy()
            g = H.download(B.irfft(H.upload(ctx, want), rdt, [1], s))
            _close(g, w, rdt, f"irfft/four-step/{shape}/s={s}")
    n = 1 << 21
    z = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    got = H.download(B.fft(H.upload(ctx, H.HostView.from_array(z, "c64")), [0]))
    want = np.fft.fft(z)
    assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max()


@pytest.mark.parametrize("rdt,cdt", [("f32", "c32"), ("f64", "c64")])
def test_rfft_irfft(ctx, oracle, rdt, cdt):
    rng = np.random.default_rng(14)
    for shape, axes in (((8,), [0]), ((9,), [0]), ((1,), [0]), ((4, 6), [0, 1]), ((4, 7), [0, 1]),
                        ((3, 5, 8), [1, 2]), ((6, 5), [1, 0]), ((2, 130), [1])):
        x = _mk(rng, shape, rdt)
        want = oracle.rfft(x, cdt, axes)
        got = H.download(B.rfft(H.upload(ctx, x), cdt, axes))
        _close(got, want.numpy(), cdt, f"rfft/{rdt}/{shape}/{axes}")
        for s in (None, [shape[a] for a in axes], [shape[a] + 3 for a in axes], [max(1, shape[a] - 2) for a in axes]):
            w = oracle.irfft(want, rdt, axes, s).numpy()
            g = H.download(B.irfft(H.upload(ctx, want), rdt, axes, s))
            _close(g, w, rdt, f"irfft/{rdt}/{shape}/{axes}/s={s}")


def test_fft_roundtrip_property_large(ctx):
    """Size-independent properties at a size no CPU oracle finishes quickly: ifft(fft x) = n x,
    linearity, and Parseval."""
    rng = np.random.default_rng(15)
    n_lines, n = 512, 4096
    a = (rng.standard_normal((n_lines, n)) + 1j * rng.standard_normal((n_lines, n))).astype(np.complex64)
    b = (rng.standard_normal((n_lines, n)) + 1j * rng.standard_normal((n_lines, n))).astype(np.complex64)
    ta, tb = H.upload(ctx, H.HostView.from_array(a, "c32")), H.upload(ctx, H.HostView.from_array(b, "c32"))
    fa, fb = H.download(B.fft(ta, [1])), H.download(B.fft(tb, [1]))
    back = H.download(B.ifft(B.fft(ta, [1]), [1])) / n
    assert np.abs(back - a).max() <= 1e-5 * max(1.0, np.abs(a).max())
    fab = H.download(B.fft(B.add(ta, tb), [1]))
    assert np.abs(fab - (fa + fb)).max() <= 1e-5 * np.abs(fab).max()
    lhs, rhs = np.sum(np.abs(fa.astype(np.complex128)) ** 2, axis=1), n * np.sum(np.abs(a.astype(np.complex128)) ** 2, axis=1)
    assert np.abs(lhs - rhs).max() <= 1e-5 * rhs.max()


def test_fft_errors(ctx):
    x = H.upload(ctx, H.HostView.from_array(np.ones((4, 4)), "f32"))
    z = H.upload(ctx, H.HostView.from_array(np.ones((4, 4)) + 0j, "c32"))
    with pytest.raises(Failure, match="fft: unsupported bigarray kind"):
        B.fft(x, [0])
    with pytest.raises(InvalidArgument, match="fft: axis out of range"):
        B.fft(z, [2])
    with pytest.raises(Failure, match="rfft: unsupported bigarray kind"):
        B.rfft(z, "c32", [0])
    with pytest.raises(Failure, match="irfft: unsupported bigarray kind"):
        B.irfft(x, "f32", [0])
