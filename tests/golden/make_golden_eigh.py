"""Generates tests/golden/nx_reference_eigh.npz from the reference's own nx_c_eigh.c (compiled
unmodified into oracle/_ref/libnxref.so):  python tests/golden/make_golden_eigh.py
Seeded Hermitian inputs (default_rng(51)); stored: the reference's eigenvalues (f64, ascending).
Eigenvectors are unique only up to a phase per column, so they are pinned by residual, not by
value (tests/test_oracle_eigh.py, tests/test_gpu_eigh.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.hostview import HostView  # noqa: E402

DTS = ("f32", "f64", "c32", "c64")
SHAPES = (((), 1), ((), 2), ((), 3), ((2,), 5), ((2, 2), 4), ((), 12), ((), 33))


def inputs():
    rng = np.random.default_rng(51)
    for dt in DTS:
        for bshape, n in SHAPES:
            a = rng.standard_normal(bshape + (n, n))
            if dt[0] == "c":
                a = a + 1j * rng.standard_normal(bshape + (n, n))
            yield f"eigh|{dt}|{bshape}|{n}", HostView.from_array(a + np.conj(np.swapaxes(a, -1, -2)), dt)


if __name__ == "__main__":
    from oracle import ref
    out = {k: ref.eigh(hv, False).numpy() for k, hv in inputs()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nx_reference_eigh.npz")
    np.savez_compressed(path, **out)
    print(f"{len(out)} vectors -> {path} ({os.path.getsize(path)} bytes)")
