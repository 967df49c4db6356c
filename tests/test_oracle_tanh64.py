"""raven_b200/csrc/nxc_tanh64.cuh (the float64 tanh the CUDA kernels use) compiled for the HOST and
compared with the host libm -- the function the reference's f64 tanh IS (nx_c_map.c:1199). The
header restates glibc's expm1-based algorithm with its fused multiply-adds explicit, so on an
FMA-capable host the two agree bit for bit; north_star's bound is 2 ulp. No GPU needed: the GPU
suite holds the kernel itself to the oracle (tests/test_gpu_map.py::test_large_unary_ulp)."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "%s/raven_b200/csrc/nxc_tanh64.cuh"
extern "C" void tanh64_many(const double *x, double *y, long n) { for (long i = 0; i < n; i++) y[i] = nxc_t64::tanh64(x[i]); }
extern "C" void expm1_many(const double *x, double *y, long n) { for (long i = 0; i < n; i++) y[i] = nxc_t64::expm1_core(x[i]); }
// the host libm itself (numpy's tanh is its own SIMD routine, not the function the reference calls)
extern "C" void libm_tanh_many(const double *x, double *y, long n) { for (long i = 0; i < n; i++) y[i] = tanh(x[i]); }
extern "C" void libm_expm1_many(const double *x, double *y, long n) { for (long i = 0; i < n; i++) y[i] = expm1(x[i]); }
''' % ROOT


def _build():
    d = tempfile.mkdtemp(prefix="nxc_t64_")
    src, so = os.path.join(d, "t.cpp"), os.path.join(d, "t.so")
    open(src, "w").write(SRC)
    # -ffp-contract=off: only the fma() calls the header spells out may fuse, as with nvcc -fmad=false
    flags = ["-O2", "-ffp-contract=off", "-shared", "-fPIC"]
    if "fma" in open("/proc/cpuinfo").read():
        flags.append("-mfma")
    subprocess.run(["g++"] + flags + ["-fno-builtin", src, "-o", so, "-lm"], check=True)
    return ctypes.CDLL(so)


def _ulps(a, b):
    ia, ib = a.view(np.int64).copy(), b.view(np.int64).copy()
    ia = np.where(ia < 0, np.int64(-2**63) - ia, ia)
    ib = np.where(ib < 0, np.int64(-2**63) - ib, ib)
    d = np.abs(ia - ib).astype(np.float64)
    return np.where(np.isnan(a) & np.isnan(b), 0.0, d)


def test_tanh64_follows_the_host_libm():
    lib = _build()
    rng = np.random.default_rng(0)
    n = 1 << 20
    x = np.concatenate([rng.uniform(-1, 1, n), rng.uniform(-4, 4, n), rng.uniform(-25, 25, n),
                        np.ldexp(rng.uniform(-1, 1, n), -rng.integers(0, 60, n)),
                        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 22.0, -22.0, 1.0, -1.0, 5e-324, 1e-300, 0.5493061443340549])])
    y = np.empty_like(x)
    ptr = ctypes.POINTER(ctypes.c_double)
    lib.tanh64_many(x.ctypes.data_as(ptr), y.ctypes.data_as(ptr), ctypes.c_long(x.size))
    want = np.empty_like(x)
    lib.libm_tanh_many(x.ctypes.data_as(ptr), want.ctypes.data_as(ptr), ctypes.c_long(x.size))
    d = _ulps(y, want)
    assert d.max() <= 2, f"worst {d.max()} ulp at x = {x[int(np.argmax(d))]!r}"
    sp = y[-12:]   # the special arguments appended last: 0, -0, inf, -inf, nan, ...
    assert sp[0] == 0 and np.signbit(sp[1]) and sp[2] == 1.0 and sp[3] == -1.0 and np.isnan(sp[4]) and sp[5] == 1.0
    # the measured agreement on an FMA host is exact; allow the non-FMA libm variant its 0.002 %
    assert (d == 0).mean() > 0.999
    xe = rng.uniform(-40, 40, n)
    ye = np.empty_like(xe)
    lib.expm1_many(xe.ctypes.data_as(ptr), ye.ctypes.data_as(ptr), ctypes.c_long(n))
    we = np.empty_like(xe)
    lib.libm_expm1_many(xe.ctypes.data_as(ptr), we.ctypes.data_as(ptr), ctypes.c_long(n))
    assert _ulps(ye, we).max() <= 1
