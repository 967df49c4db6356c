"""Generates tests/golden/nx_reference_tier3.npz from the reference's own nx_c_svd.c / nx_c_eig.c
(compiled unmodified into oracle/_ref/libnxref.so):  python tests/golden/make_golden_tier3.py
Seeded inputs (default_rng(73)); stored: the reference's singular values (f64, descending) and its
eigenvalues (c64, sorted by (real, imag) because the contract fixes no order). Singular / eigen
VECTORS are unique only up to a phase (and a basis inside a repeated value), so they are pinned by
their defining properties (tests/test_oracle_tier3.py, tests/test_gpu_tier3.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.hostview import HostView  # noqa: E402

DTS = ("f32", "f64", "c32", "c64")
SVD_SHAPES = ((1, 1), (4, 3), (3, 4), (5, 5), (2, 6, 4), (2, 2, 3, 5), (1, 7), (7, 1), (33, 17), (12, 30))
EIG_SHAPES = ((1, 1), (2, 2), (3, 3), (2, 5, 5), (2, 2, 4, 4), (12, 12), (33, 33))


def sort_eigs(w):
    """canonical order of an eigenvalue set: by real part, then imaginary (rounded so that rounding
    noise in a conjugate pair's real parts cannot swap them)"""
    flat = w.reshape((-1, w.shape[-1]))
    out = np.empty_like(flat)
    for i, r in enumerate(flat):
        scale = max(1.0, float(np.abs(r).max()))
        out[i] = r[np.lexsort((np.round(r.imag / scale, 6), np.round(r.real / scale, 6)))]
    return out.reshape(w.shape)


def svd_inputs():
    rng = np.random.default_rng(73)
    for dt in DTS:
        for shp in SVD_SHAPES:
            a = rng.standard_normal(shp)
            if dt[0] == "c":
                a = a + 1j * rng.standard_normal(shp)
            yield f"svd|{dt}|{shp}", HostView.from_array(a, dt)
        # rank-deficient: rank 1, exact zeros, a repeated singular value
        x, y = rng.standard_normal(6), rng.standard_normal(4)
        yield f"svd|{dt}|rank1", HostView.from_array(np.outer(x, y), dt)
        yield f"svd|{dt}|zeros", HostView.from_array(np.zeros((3, 5)), dt)
        yield f"svd|{dt}|eye", HostView.from_array(np.eye(5, 3) * 2.0, dt)


def eig_inputs():
    rng = np.random.default_rng(74)
    for dt in DTS:
        for shp in EIG_SHAPES:
            a = rng.standard_normal(shp)
            if dt[0] == "c":
                a = a + 1j * rng.standard_normal(shp)
            yield f"eig|{dt}|{shp}", HostView.from_array(a, dt)
        # the contract suite's matrix (backend_contract.ml:2241-2242), a rotation (purely imaginary pair),
        # a triangular matrix (eigenvalues on the diagonal)
        yield f"eig|{dt}|contract", HostView.from_array(np.array([[2., -1, 0], [1, 3, -1], [0, 1, 2]]), dt)
        yield f"eig|{dt}|rot", HostView.from_array(np.array([[0., -1], [1, 0]]), dt)
        yield f"eig|{dt}|triu", HostView.from_array(np.triu(rng.standard_normal((6, 6))), dt)


if __name__ == "__main__":
    from oracle import ref
    out = {k: ref.svd(hv, False)[1].numpy() for k, hv in svd_inputs()}
    out.update({k: sort_eigs(ref.eig(hv, False).numpy()) for k, hv in eig_inputs()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nx_reference_tier3.npz")
    np.savez_compressed(path, **out)
    print(f"{len(out)} vectors -> {path} ({os.path.getsize(path)} bytes)")
