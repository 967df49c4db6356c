"""Generates tests/golden/nx_reference_linalg.npz from the reference's own nx_c_tri.c /
nx_c_qr.c (compiled unmodified into oracle/_ref/libnxref.so by oracle/Makefile):
    python tests/golden/make_golden_linalg.py
Seeded inputs (default_rng(31)); outputs are the reference's cholesky (lower / upper),
triangular_solve (all flag combinations used by the frontend) and qr (reduced / full) results,
stored as float64 / complex128 so the checker compares values, not storage bits."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import nxo  # noqa: E402  (only its `cast`, for the 16-bit storage types)
from oracle.hostview import HostView  # noqa: E402

DTS = ("f32", "f64", "c32", "c64", "bf16", "f16")
WIDE = {"f32": "f64", "f64": "f64", "bf16": "f64", "f16": "f64", "c32": "c64", "c64": "c64"}


def hv_of(a, dt):
    if dt in ("bf16", "f16"):
        return nxo.cast(HostView.from_array(a.astype(np.float32), "f32"), dt)
    return HostView.from_array(a, dt)


def wide(hv):
    return nxo.cast(hv, WIDE[hv.dtype]).numpy()


def _mk(rng, shape, dt):
    a = rng.standard_normal(shape)
    if dt in ("c32", "c64"):
        a = a + 1j * rng.standard_normal(shape)
    return a


def cases():
    rng = np.random.default_rng(31)
    for dt in DTS:
        for bshape, n in (((), 1), ((), 4), ((2,), 7), ((2, 3), 5), ((), 40)):
            a = _mk(rng, bshape + (n, n), dt)
            spd = hv_of(a @ np.conj(np.swapaxes(a, -1, -2)) + n * np.eye(n), dt)
            for upper in (False, True):
                yield f"chol|{dt}|{bshape}|{n}|{int(upper)}", (lambda m, spd=spd, upper=upper: wide(m.cholesky(spd, upper)))
            tri = {False: hv_of(np.tril(a) / n + np.eye(n), dt), True: hv_of(np.triu(a) / n + np.eye(n), dt)}
            for nrhs in (1, 3):
                b = hv_of(_mk(rng, bshape + (n, nrhs), dt), dt)
                for upper, tr, unit in ((False, False, False), (False, True, False), (True, False, True),
                                        (True, True, False), (False, False, True)):
                    yield (f"trsm|{dt}|{bshape}|{n}|{nrhs}|{int(upper)}{int(tr)}{int(unit)}",
                           (lambda m, A=tri[upper], b=b, f=(upper, tr, unit): wide(m.triangular_solve(A, b, *f))))
        for bshape, mm, nn in (((), 4, 4), ((2,), 6, 3), ((), 3, 6), ((2,), 1, 1), ((), 30, 20)):
            x = hv_of(_mk(rng, bshape + (mm, nn), dt), dt)
            for red in (True, False):
                yield f"qr_q|{dt}|{bshape}|{mm}x{nn}|{int(red)}", (lambda m, x=x, red=red: wide(m.qr(x, red)[0]))
                yield f"qr_r|{dt}|{bshape}|{mm}x{nn}|{int(red)}", (lambda m, x=x, red=red: wide(m.qr(x, red)[1]))


if __name__ == "__main__":
    from oracle import ref
    out = {k: t(ref) for k, t in cases()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nx_reference_linalg.npz")
    np.savez_compressed(path, **out)
    print(f"{len(out)} vectors -> {path} ({os.path.getsize(path)} bytes)")
