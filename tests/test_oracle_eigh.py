"""The eigh restatement (oracle/nxo.py, cyclic Jacobi in double precision) against the eigenvalues
the reference's own nx_c_eigh.c produced (tests/golden/nx_reference_eigh.npz) and, for the
eigenvectors, against the defining properties (A v = w v, orthonormal columns): they are unique
only up to a phase per column. Tolerances relative to the spectral radius: the input type's
rounding (the reference computes f32 inputs in f32)."""
import os

import numpy as np
import pytest

from oracle import nxo, ref
from tests.golden.make_golden_eigh import inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nx_reference_eigh.npz")
TOL = {"f32": 2e-5, "f64": 1e-12, "c32": 2e-5, "c64": 1e-12}


def check_eigh(w, v, hv, dt, key, tol_scale=1.0):
    a = hv.numpy().astype(np.complex128)
    n = a.shape[-1]
    low = np.tril(a)
    a = low + np.conj(np.swapaxes(np.tril(a, -1), -1, -2))
    scale = max(1.0, float(np.abs(w).max()))
    assert np.all(np.diff(w, axis=-1) >= 0), key
    V = v.astype(np.complex128)
    assert np.abs(a @ V - V * w[..., None, :]).max() <= 50 * n * TOL[dt] * scale * tol_scale, key
    eye = np.conj(np.swapaxes(V, -1, -2)) @ V
    assert np.abs(eye - np.eye(n)).max() <= 50 * n * TOL[dt] * tol_scale, key


def test_eigh_restatement_matches_reference_golden_eigenvalues():
    gold = np.load(GOLD)
    n = 0
    for key, hv in inputs():
        dt = key.split("|")[1]
        w, v = nxo.eigh(hv)
        want = gold[key]
        assert np.abs(w.numpy() - want).max() <= 10 * TOL[dt] * max(1.0, np.abs(want).max()), key
        check_eigh(w.numpy(), v.numpy(), hv, dt, key)
        assert np.array_equal(nxo.eigh(hv, False).numpy(), w.numpy())
        n += 1
    assert n == len(gold.files) == 28


@pytest.mark.skipif(not ref.available(), reason="reference binary not built")
def test_reference_eigh_meets_the_same_properties():
    for key, hv in inputs():
        w, v = ref.eigh(hv)
        check_eigh(w.numpy(), v.numpy(), hv, key.split("|")[1], key)
