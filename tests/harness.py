"""Shared machinery of the parity tests.

The same host bytes feed the checker (oracle/) and the CUDA path: a HostView is
uploaded verbatim and the device handle gets the HostView's shape / strides /
offset, so strided, broadcast, flipped and offset views are exercised on both
sides without any materialisation in between.

Pools and the layout matrix are the reference contract suite's
(packages/nx/test/backend_contract.ml:125-289, 423-468): the float pool holds
0.0 and negatives so log/sqrt/asin/acos/recip hit NaN/inf; unsigned pools set
the high bit so signed/unsigned interpretations diverge.
"""
from __future__ import annotations

import numpy as np

from oracle.hostview import HostView, DTYPES, FLOATS, SINTS, UINTS, INTS, COMPLEX, np_storage  # noqa: F401

FPOOL = [0.5, -1.5, 2.0, -0.25, 3.0, 1.0, -2.0, 0.75, -3.5, 4.0, -0.5, 2.5, 1.25, -1.0, 0.0, 5.0, -4.0, 1.75]
IPOOL_S = [3, -7, 1, -10, 5, 2, -8, 4, 9, -6, 0, 12, -11, 15, -13, 14, -20, 17]
IPOOL_U = [3, 200, 1, 250, 5, 2, 130, 4, 9, 6, 0, 255, 11, 128, 13, 14, 240, 17]
U16POOL = [3, 40000, 1, 250, 5, 2, 32768, 4, 9, 6, 0, 60000, 11, 128, 13, 14, 50000, 17]
U32POOL = [3, 0x80000000, 1, 0xFFFFFFF0, 5, 2, 0xC0000000, 4, 9, 6, 0, 0x90000000, 11, 0xA5A5A5A5, 13, 14,
           0xF0000000, 17]
U64POOL = [3, 0x8000000000000000, 1, 0xFFFFFFFFFFFFFFF0, 5, 2, 0xC000000000000000, 4, 9, 6, 0,
           0x9000000000000000, 11, 0xA5A5A5A5A5A5A5A5, 13, 14, 0xF000000000000000, 17]


# ---- float <-> storage bit helpers (numpy only; used to BUILD inputs, never to check) ----
def f32_to_bf16_bits(x):
    b = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((b + 0x7FFF + ((b >> 16) & 1)) >> 16).astype(np.uint16)
    return r


def bf16_bits_to_f32(b):
    return (np.asarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def to_storage(dtype: str, values) -> np.ndarray:
    """Encode python/numpy numbers as the dtype's storage array (exact for the pools)."""
    v = np.asarray(values)
    if dtype == "f16":
        return v.astype(np.float16).view(np.uint16)
    if dtype == "bf16":
        return f32_to_bf16_bits(v.astype(np.float32))
    if dtype in ("f8e4m3", "f8e5m2"):
        return np.array([encode_fp8(dtype, float(x)) for x in v.reshape(-1)], dtype=np.uint8).reshape(v.shape)
    if dtype == "bool":
        return (v != 0).astype(np.uint8)
    if dtype in ("u32", "u64", "u8", "u16", "i8", "i16", "i32", "i64"):
        return np.array([int(x) for x in v.reshape(-1)], dtype=object).astype(np_storage(dtype)).reshape(v.shape) \
            if v.dtype == object else v.astype(np_storage(dtype))
    return v.astype(np_storage(dtype))


def encode_fp8(dtype, x: float) -> int:
    """Nearest-even encode for building fp8 test inputs (exact on grid points)."""
    table = fp8_table(dtype)
    if x != x:
        return 0x7F
    finite = [(abs(table[c] - x), c) for c in range(256) if table[c] == table[c] and abs(table[c]) != float("inf")]
    best = min(finite)[0]
    cands = [c for d, c in finite if d == best]
    cands.sort(key=lambda c: (c & 1, c))
    return cands[0]


_FP8 = {}


def fp8_table(dtype):
    if dtype not in _FP8:
        t = []
        for c in range(256):
            s = -1.0 if c & 0x80 else 1.0
            if dtype == "f8e4m3":
                e, m = (c >> 3) & 0xF, c & 7
                if e == 0xF and m == 7:
                    t.append(float("nan"))
                elif e == 0:
                    t.append(s * m / 8.0 * 2.0 ** -6)
                else:
                    t.append(s * (1 + m / 8.0) * 2.0 ** (e - 7))
            else:
                e, m = (c >> 2) & 0x1F, c & 3
                if e == 0x1F:
                    t.append(s * float("inf") if m == 0 else float("nan"))
                elif e == 0:
                    t.append(s * m / 4.0 * 2.0 ** -14)
                else:
                    t.append(s * (1 + m / 4.0) * 2.0 ** (e - 15))
        _FP8[dtype] = np.array(t, dtype=np.float64)
    return _FP8[dtype]


def storage_to_float(dtype: str, s: np.ndarray) -> np.ndarray:
    """Decode storage to float64 for tolerance comparisons of low-precision floats."""
    if dtype == "f16":
        return s.view(np.float16).astype(np.float64)
    if dtype == "bf16":
        return bf16_bits_to_f32(s).astype(np.float64)
    if dtype in ("f8e4m3", "f8e5m2"):
        return fp8_table(dtype)[s]
    return s.astype(np.float64)


def pool(dtype: str, n: int = 18) -> np.ndarray:
    if dtype in FLOATS:
        base = FPOOL
    elif dtype in SINTS:
        base = IPOOL_S
    elif dtype == "u8":
        base = IPOOL_U
    elif dtype == "u16":
        base = U16POOL
    elif dtype == "u32":
        base = U32POOL
    elif dtype == "u64":
        base = U64POOL
    elif dtype == "bool":
        base = [1, 0, 1, 1, 0, 0, 1, 0, 1, 1, 1, 0, 0, 1, 0, 1, 0, 0]
    elif dtype in COMPLEX:
        base = [complex(FPOOL[i], FPOOL[(i + 5) % 18]) for i in range(18)]
    else:
        raise KeyError(dtype)
    reps = (n + len(base) - 1) // len(base)
    vals = (base * reps)[:n]
    if dtype in ("u32", "u64"):
        return np.array(vals, dtype=np_storage(dtype))
    return to_storage(dtype, np.array(vals))


def layouts(dtype: str, rot: int = 0, include_degenerate: bool = True):
    """The contract's layout matrix: (name, HostView). `rot` rotates the pool so two
    operands of a binary op see different values."""
    p = np.roll(pool(dtype, 18), -rot)
    hv = lambda shape, n: HostView(p[:n].copy(), dtype, shape)  # noqa: E731
    out = [
        ("contig", hv([3, 4], 12)),
        ("transpose", hv([4, 3], 12).permute([1, 0])),
        ("slice", hv([3, 6], 18).shrink([(0, 3), (1, 5)])),
        ("broadcast", hv([1, 4], 4).expand([3, 4])),
        ("flip", hv([3, 4], 12).flip([1])),
        ("permute3", hv([2, 3, 2], 12).permute([2, 0, 1])),
        ("rank5", hv([1, 3, 1, 2, 2], 12)),
    ]
    if include_degenerate:
        out += [("scalar", hv([], 1)), ("empty", HostView(np.zeros(0, dtype=np_storage(dtype)), dtype, [0, 4]))]
    return out


# binary ops vary both operand layouts independently (backend_contract.ml:555-593)
BINARY_LAYOUT_PAIRS = [("contig", "contig"), ("contig", "transpose"), ("broadcast", "contig"),
                       ("flip", "slice"), ("scalar", "scalar"), ("empty", "empty"), ("rank5", "rank5"),
                       ("permute3", "permute3")]


# ---- oracle selection -----------------------------------------------------------------
def get_oracle():
    """The reference's own C (oracle/_ref) when it has been built, else the C
    restatement (oracle/nxo). Both expose the same functions."""
    from oracle import ref
    if ref.available():
        return ref
    from oracle import nxo
    return nxo


# ---- device transfer ------------------------------------------------------------------------
def upload(ctx, hv: HostView):
    import raven_b200.backend as B
    from raven_b200 import dtype as D
    base = B.from_host(ctx, hv.storage, D.of(hv.dtype))
    return B.Tensor(base.buffer, hv.shape, hv.strides, hv.offset, base.dtype, ctx)


def download(t) -> np.ndarray:
    import raven_b200.backend as B
    return B.to_numpy(t)


def raw(x: np.ndarray) -> np.ndarray:
    """The bytes of an array (works for 0-d too)."""
    return np.ascontiguousarray(x).reshape(-1).view(np.uint8)


# ---- comparison ------------------------------------------------------------------------------
def ulp_diff(dtype: str, got: np.ndarray, want: np.ndarray) -> np.ndarray:
    """Distance in units in the last place of the STORAGE type; NaN vs NaN is 0,
    NaN vs number is huge; +0 and -0 are 0 apart."""
    if dtype in ("f32", "f64"):
        it = np.int32 if dtype == "f32" else np.int64
        g = got.view(it).astype(np.int64) if dtype == "f32" else got.view(it)
        w = want.view(it).astype(np.int64) if dtype == "f32" else want.view(it)
        sign = np.int64(np.iinfo(it).min)
        g = np.where(g < 0, sign - g, g)
        w = np.where(w < 0, sign - w, w)
        with np.errstate(over="ignore"):
            d = np.abs(g - w).astype(np.float64)  # integer domain: exact for neighbouring values
    else:
        bits = {"f16": (np.int16, 16), "bf16": (np.int16, 16), "f8e4m3": (np.int8, 8), "f8e5m2": (np.int8, 8)}[dtype]
        g = got.view(bits[0]).astype(np.int64)
        w = want.view(bits[0]).astype(np.int64)
        sign = -(1 << (bits[1] - 1))
        g = np.where(g < 0, sign - g, g)
        w = np.where(w < 0, sign - w, w)
        d = np.abs(g - w).astype(np.float64)
    gn = np.isnan(storage_to_float(dtype, got))
    wn = np.isnan(storage_to_float(dtype, want))
    d = np.where(gn & wn, 0.0, d)
    d = np.where(gn ^ wn, np.inf, d)
    return d


def assert_same(dtype: str, got: np.ndarray, want: np.ndarray, ulp: float = 0, what: str = ""):
    """Bit-exact for ints/bool (and floats when ulp == 0, modulo NaN payload and
    the sign of zero); within `ulp` storage ulps for floats; per component for complex."""
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    if got.size == 0:
        return
    if dtype in INTS or dtype == "bool" or dtype in ("i4", "u4"):
        if not np.array_equal(got, want):
            bad = np.argwhere(got != want)[:5]
            raise AssertionError(f"{what}: {dtype} mismatch at {bad.tolist()}: got {got[tuple(bad[0])]} want {want[tuple(bad[0])]}")
        return
    if dtype in COMPLEX:
        part = "f32" if dtype == "c32" else "f64"
        rt = np.float32 if dtype == "c32" else np.float64
        g = np.ascontiguousarray(got).reshape(-1).view(rt)
        w = np.ascontiguousarray(want).reshape(-1).view(rt)
        # error relative to the modulus: a tiny component next to a large one carries the large one's ulp
        with np.errstate(all="ignore"):
            mod = np.repeat(np.abs(want).reshape(-1), 2).astype(np.float64)
        mod = np.where(np.isfinite(mod), mod, 0.0)
        eps = np.finfo(rt).eps
        err = np.abs(g.astype(np.float64) - w.astype(np.float64))
        ok = (err <= max(ulp, 1) * eps * np.maximum(mod, np.finfo(rt).tiny)) | (np.isnan(g) & np.isnan(w)) | (g == w)
        assert ok.all(), f"{what}: {dtype} mismatch: got {g[~ok][:4]} want {w[~ok][:4]}"
        return
    d = ulp_diff(dtype, got, want)
    if not (d <= ulp).all():
        i = np.unravel_index(np.argmax(d), d.shape)
        raise AssertionError(f"{what}: {dtype} max ulp diff {d.max()} > {ulp} at {i}: got "
                             f"{storage_to_float(dtype, got)[i]!r} want {storage_to_float(dtype, want)[i]!r}")


def assert_close(dtype: str, got: np.ndarray, want: np.ndarray, rel: float, abs_: float = 0.0, what: str = ""):
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    if dtype in COMPLEX:
        g, w = got.astype(np.complex128), want.astype(np.complex128)
    else:
        g, w = storage_to_float(dtype, got), storage_to_float(dtype, want)
    both_nan = np.isnan(g) & np.isnan(w)
    same_inf = np.isinf(g) & np.isinf(w) & (g == w)
    err = np.abs(np.where(both_nan | same_inf, 0, g - w))
    tol = abs_ + rel * np.abs(np.where(both_nan | same_inf, 0, w))
    ok = (err <= tol) | both_nan | same_inf
    assert ok.all(), f"{what}: {dtype} got {g[~ok][:4]} want {w[~ok][:4]} (rel {rel})"


def storage_ulp(dtype, x):
    """Spacing of the 16-bit storage type at |x| (float64 array)."""
    mant, emin = {"bf16": (7, -126), "f16": (10, -14)}[dtype]
    ax = np.abs(np.asarray(x, dtype=np.float64))
    e = np.floor(np.log2(np.where(ax > 0, ax, 1.0)))
    e = np.where(ax > 0, np.maximum(e, emin), emin)
    return np.exp2(e - mant)


def assert_gemm_16bit(dtype, got, want, absprod, k, what):
    """bf16 / f16 products, ELEMENTWISE: the reference accumulates in f32 and rounds once
    (nx_c_matmul.c:12-28); so does the tensor-core path, in another order. Two f32 sums of the same
    K products differ by at most ~2 K eps32 sum|a||b| (which matters only where the sum cancels),
    and the single rounding to storage can then land on the neighbouring value: 1 storage ulp of the
    reference's element plus that accumulation slack, per element -- no tolerance relative to the
    largest element of C."""
    g, w = storage_to_float(dtype, got), storage_to_float(dtype, want)
    bound = storage_ulp(dtype, w) + 2.0 * k * 2.0 ** -24 * absprod
    err = np.abs(g - w)
    ok = (err <= bound) | (np.isnan(g) & np.isnan(w)) | (g == w)
    if not ok.all():
        i = np.unravel_index(np.argmax(np.where(ok, 0, err / bound)), err.shape)
        raise AssertionError(f"{what}: {dtype} element {i}: got {g[i]!r} want {w[i]!r}, |diff| {err[i]:.3e} > 1 ulp + "
                             f"slack = {bound[i]:.3e} ({int((~ok).sum())} of {ok.size} elements)")
