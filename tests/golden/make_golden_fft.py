"""Generates tests/golden/nx_reference_fft.npz from the reference's own nx_c_fft.c (compiled
unmodified into oracle/_ref/libnxref.so by oracle/Makefile):
    python tests/golden/make_golden_fft.py
Inputs are seeded (default_rng(21), standard normal); outputs are the reference's results for
fft / ifft / rfft / irfft over one and several axes, power-of-two, smooth and prime lengths,
and explicit irfft sizes. The checker (tests/test_oracle_fft.py) pins the direct-DFT
restatement (oracle/nxo.py) to them where /root/reference does not exist."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.hostview import HostView  # noqa: E402

C_CASES = [((1,), [0]), ((2,), [0]), ((8,), [0]), ((12,), [0]), ((17,), [0]), ((100,), [0]), ((64,), [0]),
           ((4, 6), [0, 1]), ((3, 5, 8), [2]), ((3, 5, 8), [0, 2]), ((5, 7), [1, 0]), ((26,), [0])]
R_CASES = [((8,), [0]), ((9,), [0]), ((1,), [0]), ((4, 6), [0, 1]), ((4, 7), [0, 1]), ((3, 5, 8), [1, 2]), ((6, 5), [1, 0])]


def _mk(rng, shape, dt):
    if dt in ("c32", "c64"):
        return HostView.from_array(rng.standard_normal(shape) + 1j * rng.standard_normal(shape), dt)
    return HostView.from_array(rng.standard_normal(shape), dt)


def cases():
    """Yields (key, thunk(module) -> numpy array); inputs are rebuilt from the seed every time."""
    rng = np.random.default_rng(21)
    for dt in ("c32", "c64"):
        for shape, axes in C_CASES:
            x = _mk(rng, shape, dt)
            for inv in (False, True):
                yield f"fft|{dt}|{shape}|{axes}|{int(inv)}", (lambda m, x=x, axes=axes, inv=inv: m.fft(x, axes, inv).numpy())
    for rdt, cdt in (("f32", "c32"), ("f64", "c64")):
        for shape, axes in R_CASES:
            x = _mk(rng, shape, rdt)
            yield f"rfft|{rdt}|{shape}|{axes}", (lambda m, x=x, cdt=cdt, axes=axes: m.rfft(x, cdt, axes).numpy())
            hshape = list(shape)
            hshape[axes[-1]] = shape[axes[-1]] // 2 + 1
            X = _mk(rng, tuple(hshape), cdt)
            for s in (None, [shape[a] for a in axes], [shape[a] + 3 for a in axes], [max(1, shape[a] - 2) for a in axes]):
                yield (f"irfft|{cdt}|{shape}|{axes}|{s}",
                       (lambda m, X=X, rdt=rdt, axes=axes, s=s: m.irfft(X, rdt, axes, s).numpy()))


if __name__ == "__main__":
    from oracle import ref
    out = {k: t(ref) for k, t in cases()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nx_reference_fft.npz")
    np.savez_compressed(path, **out)
    print(f"{len(out)} vectors -> {path} ({os.path.getsize(path)} bytes)")
