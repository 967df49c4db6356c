"""linalg tier 3 on the CPU: (1) the svd / eig restatements (oracle/nxo.py) against what the
reference's own nx_c_svd.c / nx_c_eig.c produced (tests/golden/nx_reference_tier3.npz: singular
values and eigenvalue sets) and against the defining properties for the vectors; (2) the KERNEL
BODIES of raven_b200/csrc/nxc_linalg3.cuh, compiled for the host and run single-threaded
(tests/emu/la3_emu.cpp -- test infrastructure, not a product path), against the same vectors.
Tolerances are relative to the largest singular / eigen value, in the input type's rounding."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import nxo, ref
from tests.golden.make_golden_tier3 import eig_inputs, sort_eigs, svd_inputs

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "nx_reference_tier3.npz")
TOL = {"f32": 2e-6, "f64": 1e-13, "c32": 2e-6, "c64": 1e-13}


def check_svd(u, s, vh, hv, dt, key, full, tol_scale=1.0):
    a = hv.numpy().astype(np.complex128)
    m, n = a.shape[-2:]
    k = min(m, n)
    assert s.shape == a.shape[:-2] + (k,) and s.dtype == np.float64, key
    assert u.shape == a.shape[:-2] + (m, m if full else k), key
    assert vh.shape == a.shape[:-2] + (n if full else k, n), key
    assert np.all(s >= 0) and np.all(np.diff(s, axis=-1) <= 0), key
    U, Vh = u.astype(np.complex128), vh.astype(np.complex128)
    tol = 50 * max(m, n) * TOL[dt] * tol_scale
    scale = max(1.0, float(s.max()))
    rec = (U[..., :, :k] * s[..., None, :]) @ Vh[..., :k, :]
    assert np.abs(rec - a).max() <= tol * scale, key
    assert np.abs(np.conj(np.swapaxes(U, -1, -2)) @ U - np.eye(U.shape[-1])).max() <= tol, key
    assert np.abs(Vh @ np.conj(np.swapaxes(Vh, -1, -2)) - np.eye(Vh.shape[-2])).max() <= tol, key


def check_eig(w, v, hv, key, tol=1e-10):
    a = hv.numpy().astype(np.complex128)
    n = a.shape[-1]
    assert w.dtype == np.complex128 and w.shape == a.shape[:-1], key
    scale = max(1.0, float(np.abs(a).max())) * n
    if v is not None:
        assert v.dtype == np.complex128 and v.shape == a.shape, key
        assert np.abs(a @ v - v * w[..., None, :]).max() <= tol * scale, key
        assert np.abs(np.linalg.norm(v, axis=-2) - 1).max() <= 1e-12, key


def eig_set_err(w, want):
    """distance between eigenvalue SETS per matrix (order-free; conjugate pairs and clusters make
    a sorted compare fragile)"""
    fw, fg = w.reshape((-1, w.shape[-1])), want.reshape((-1, want.shape[-1]))
    err = 0.0
    for a, b in zip(fw, fg):
        d = np.abs(a[:, None] - b[None, :])
        err = max(err, d.min(axis=0).max(), d.min(axis=1).max())
    return err


def test_golden_inventory():
    gold = np.load(GOLD)
    assert len(gold.files) == 92


def test_svd_restatement_matches_reference_golden():
    gold = np.load(GOLD)
    for key, hv in svd_inputs():
        dt = key.split("|")[1]
        for full in (False, True):
            u, s, vh = nxo.svd(hv, full)
            want = gold[key]
            assert np.abs(s.numpy() - want).max() <= 20 * TOL[dt] * max(1.0, want.max()), key
            check_svd(u.numpy(), s.numpy(), vh.numpy(), hv, dt, key, full)


def test_eig_restatement_matches_reference_golden():
    gold = np.load(GOLD)
    for key, hv in eig_inputs():
        w, v = nxo.eig(hv)
        want = gold[key]
        # eigenvalues of a nonnormal matrix move by eps * cond: 1e-9 of the spectral radius for these sizes
        assert eig_set_err(w.numpy(), want) <= 1e-9 * max(1.0, np.abs(want).max()), key
        check_eig(w.numpy(), v.numpy(), hv, key)
        assert np.array_equal(nxo.eig(hv, False).numpy(), w.numpy()), key


@pytest.mark.skipif(not ref.available(), reason="reference binary not built")
def test_reference_meets_the_same_properties():
    gold = np.load(GOLD)
    for key, hv in svd_inputs():
        for full in (False, True):
            u, s, vh = ref.svd(hv, full)
            assert np.array_equal(s.numpy(), gold[key]), key
            check_svd(u.numpy(), s.numpy(), vh.numpy(), hv, key.split("|")[1], key, full)
    for key, hv in eig_inputs():
        w, v = ref.eig(hv)
        assert np.array_equal(sort_eigs(w.numpy()), gold[key]), key
        if key.split("|")[2] != "triu":  # (repeated-root free inputs only: a residual check is then meaningful)
            check_eig(w.numpy(), v.numpy(), hv, key)


# ---- the kernel bodies, emulated -----------------------------------------------------------------
@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "emu", "la3_emu.cpp")
    out = os.path.join(HERE, "emu", "libla3_emu.so")
    hdr = os.path.join(os.path.dirname(HERE), "raven_b200", "csrc", "nxc_linalg3.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    lib.la3_emu_svd.argtypes = [ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int64] * 4 + [ctypes.c_void_p] * 3
    lib.la3_emu_eig.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return lib


NP = {"f32": np.float32, "f64": np.float64, "c32": np.complex64, "c64": np.complex128}


def emu_svd(lib, a, full):
    dt = {np.dtype(v): i for i, v in enumerate((np.float32, np.float64, np.complex64, np.complex128))}[a.dtype]
    m, n = a.shape[-2:]
    k = min(m, n)
    fa = np.ascontiguousarray(a).reshape((-1, m, n))
    uc, vr = (m, n) if full else (k, k)
    U, S, Vh = np.zeros((len(fa), m, uc), a.dtype), np.zeros((len(fa), k)), np.zeros((len(fa), vr, n), a.dtype)
    for b in range(len(fa)):
        assert lib.la3_emu_svd(dt, fa[b].ctypes.data, m, n, uc, vr, U[b].ctypes.data, S[b].ctypes.data, Vh[b].ctypes.data) == 0
    bs = a.shape[:-2]
    return U.reshape(bs + (m, uc)), S.reshape(bs + (k,)), Vh.reshape(bs + (vr, n))


def emu_eig(lib, a, vectors=True):
    n = a.shape[-1]
    fa = np.ascontiguousarray(a.astype(np.complex128)).reshape((-1, n, n))
    W, V = np.zeros((len(fa), n), np.complex128), np.zeros((len(fa), n, n), np.complex128)
    for b in range(len(fa)):
        assert lib.la3_emu_eig(fa[b].ctypes.data, n, int(vectors), W[b].ctypes.data, V[b].ctypes.data) == 0
    return W.reshape(a.shape[:-1]), V.reshape(a.shape)


def test_svd_kernel_body_emulated_matches_reference_golden(emu):
    gold = np.load(GOLD)
    for key, hv in svd_inputs():
        dt = key.split("|")[1]
        for full in (False, True):
            u, s, vh = emu_svd(emu, hv.numpy().astype(NP[dt]), full)
            want = gold[key]
            assert np.abs(s - want).max() <= 20 * TOL[dt] * max(1.0, want.max()), key
            check_svd(u, s, vh, hv, dt, key, full)


def test_svd_kernel_body_mixed_output_shapes(emu):
    """thin U with full V^H and the reverse: each output's shape is read on its own (nx_c_svd.c:2752-2757)"""
    a = np.random.default_rng(5).standard_normal((3, 6))
    m, n, k = 3, 6, 3
    for uc, vr in ((k, n), (m, k)):
        U, S, Vh = np.zeros((m, uc)), np.zeros(k), np.zeros((vr, n))
        assert emu.la3_emu_svd(1, a.ctypes.data, m, n, uc, vr, U.ctypes.data, S.ctypes.data, Vh.ctypes.data) == 0
        assert np.abs((U[:, :k] * S) @ Vh[:k] - a).max() <= 1e-13
        assert np.abs(Vh @ Vh.T - np.eye(vr)).max() <= 1e-13


def test_eig_kernel_body_emulated_matches_reference_golden(emu):
    gold = np.load(GOLD)
    for key, hv in eig_inputs():
        dt = key.split("|")[1]
        w, v = emu_eig(emu, hv.numpy().astype(NP[dt]))
        want = gold[key]
        assert eig_set_err(w, want) <= 1e-9 * max(1.0, np.abs(want).max()), key
        check_eig(w, v, hv, key)
        w2, _ = emu_eig(emu, hv.numpy().astype(NP[dt]), vectors=False)
        assert np.array_equal(w, w2), key


def test_eig_kernel_body_defective_matrix_stays_finite(emu):
    """a Jordan-like matrix: correct eigenvalues, finite unit eigenvectors, valid eigenpairs
    (nx_c_eig.c:55-58)"""
    a = np.triu(np.ones((60, 60)))
    w, v = emu_eig(emu, a)
    assert np.abs(w - 1).max() <= 1e-12 and np.isfinite(v).all()
    assert np.abs(a @ v - v * w[None, :]).max() <= 1e-10


def _badly_scaled(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, n))
    d = 10.0 ** rng.integers(-6, 7, n)
    return a, (a * d[None, :]) / d[:, None]   # D^-1 A D: the same spectrum, entries spread over 24 decades


def test_eig_kernel_body_balances_badly_scaled_input(emu):
    """the scaling half of the reference's `balanc` (nx_c_eig.c:25-27): without it the QR iteration's
    eps * ||A|| errors swamp the small eigenvalues of D^-1 A D"""
    for n, seed in ((8, 1), (20, 2), (50, 3)):
        a, bad = _badly_scaled(n, seed)
        want = np.linalg.eigvals(a)
        w, v = emu_eig(emu, bad)
        assert eig_set_err(w, want) <= 1e-12 * np.abs(want).max(), n
        assert np.abs(bad @ v - v * w[None, :]).max() <= 1e-10 * np.abs(bad).max()
        assert np.abs(np.linalg.norm(v, axis=0) - 1).max() <= 1e-12
