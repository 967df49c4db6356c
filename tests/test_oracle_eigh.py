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


def test_eigh_kernel_body_emulated_matches_reference_golden_eigenvalues():
    """the GPU kernel's body (one-sided Jacobi, raven_b200/csrc/nxc_linalg3.cuh), compiled for the
    host and run single-threaded (tests/emu: test infrastructure, not a product path)"""
    import ctypes
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    src, out = os.path.join(here, "emu", "la3_emu.cpp"), os.path.join(here, "emu", "libla3_emu.so")
    hdr = os.path.join(os.path.dirname(here), "raven_b200", "csrc", "nxc_linalg3.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    lib.la3_emu_eigh.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    npdt = {"f32": np.float32, "f64": np.float64, "c32": np.complex64, "c64": np.complex128}
    gold = np.load(GOLD)
    for key, hv in inputs():
        dt = key.split("|")[1]
        a = np.ascontiguousarray(hv.numpy().astype(npdt[dt]))
        n = a.shape[-1]
        fa = a.reshape((-1, n, n))
        W, V = np.zeros((len(fa), n)), np.zeros(fa.shape, npdt[dt])
        for b in range(len(fa)):
            junk = np.tril(fa[b]) + np.triu(np.full((n, n), 7.0, npdt[dt]), 1)   # only the lower triangle is read
            junk = np.ascontiguousarray(junk)
            assert lib.la3_emu_eigh(list(npdt).index(dt), junk.ctypes.data, n, 1, W[b].ctypes.data, V[b].ctypes.data) == 0
            w2 = np.zeros(n)
            assert lib.la3_emu_eigh(list(npdt).index(dt), junk.ctypes.data, n, 0, w2.ctypes.data, None) == 0
            assert np.array_equal(w2, W[b]), key
        w, v = W.reshape(a.shape[:-1]), V.reshape(a.shape)
        want = gold[key]
        assert np.abs(w - want).max() <= 10 * TOL[dt] * max(1.0, np.abs(want).max()), key
        check_eigh(w, v, hv, dt, key)
