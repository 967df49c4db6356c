"""GPU parity of the map family (unary / binary / compare / where / cast / copy /
assign / fill) against the oracle, through the C ABI, over the reference
contract's layout matrix (backend_contract.ml:423-468, 555-593) and on large
seeded arrays.

Tolerances (north_star): integer, bool, compare, where, copy, cast-to-int:
bit-exact. Float arithmetic that is a single IEEE operation (neg abs sign add sub
mul fdiv recip sqrt idiv mod max min rounding): bit-exact. Transcendentals:
<= 2 ulp for f32 / f64 against the reference's libm; <= 1 ulp of the storage type
for f16 / bf16 / fp8 (they compute in f32 and round once).
"""
import numpy as np
import pytest

import raven_b200.backend as B
from raven_b200 import Failure, InvalidArgument
from tests import harness as H

pytestmark = pytest.mark.gpu

UNARY = B.UNARY_OPS
EXACT_UNARY = {"neg", "recip", "abs", "sign", "sqrt", "trunc", "ceil", "floor", "round"}
BINARY = "add sub mul idiv fdiv mod max min pow atan2 xor or and shl shr".split()
EXACT_BINARY = {"add", "sub", "mul", "idiv", "fdiv", "mod", "max", "min", "xor", "or", "and", "shl", "shr"}
ALL = list(H.FLOATS) + list(H.INTS) + list(H.COMPLEX) + ["bool"]
BFN = {"mod": "mod_", "or": "or_", "and": "and_"}


def _ulp(dtype, exact):
    if dtype in H.INTS or dtype == "bool":
        return 0
    if exact:
        return 0
    if dtype in ("f32", "f64"):
        return 2
    return 1


def _run_both(ofn, gfn):
    """Returns ('ok', oracle_result, gpu_result) or ('err', kind, msg) after checking
    both sides fail the same way."""
    try:
        want = ofn()
    except Exception as e:  # oracle raised: product must raise the same class + message
        kind, msg = e.kind, e.msg
        with pytest.raises(InvalidArgument if kind == "Invalid_argument" else Failure) as ei:
            gfn()
        assert str(ei.value).startswith(msg), (str(ei.value), msg)
        return ("err", kind, msg)
    return ("ok", want, gfn())


@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("op", UNARY)
def test_unary_layouts(ctx, oracle, op, dtype):
    for name, hv in H.layouts(dtype):
        r = _run_both(lambda: oracle.unary(op, hv), lambda: getattr(B, op)(H.upload(ctx, hv)))
        if r[0] == "err":
            return
        if dtype in H.COMPLEX and op not in ("neg", "abs"):
            # complex: the contract's tolerance (backend_contract.ml:2329-2382), relative to the modulus
            tol = 1e-5 if dtype == "c32" else 1e-11
            H.assert_close(dtype, H.download(r[2]), r[1].numpy(), rel=tol, abs_=tol, what=f"{op}/{dtype}/{name}")
            continue
        H.assert_same(dtype, H.download(r[2]), r[1].numpy(), ulp=_ulp(dtype, op in EXACT_UNARY),
                      what=f"{op}/{dtype}/{name}")


@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("op", BINARY)
def test_binary_layouts(ctx, oracle, op, dtype):
    la = dict(H.layouts(dtype, rot=0))
    lb = dict(H.layouts(dtype, rot=5))
    for na, nb in H.BINARY_LAYOUT_PAIRS:
        a, b = la[na], lb[nb]
        r = _run_both(lambda: oracle.binary(op, a, b),
                      lambda: getattr(B, BFN.get(op, op))(H.upload(ctx, a), H.upload(ctx, b)))
        if r[0] == "err":
            return
        if dtype in H.COMPLEX and op not in ("add", "sub"):
            tol = 1e-5 if dtype == "c32" else 1e-11
            H.assert_close(dtype, H.download(r[2]), r[1].numpy(), rel=tol, abs_=tol, what=f"{op}/{dtype}/{na},{nb}")
            continue
        H.assert_same(dtype, H.download(r[2]), r[1].numpy(), ulp=_ulp(dtype, op in EXACT_BINARY),
                      what=f"{op}/{dtype}/{na},{nb}")


@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("op", B.CMP_OPS)
def test_compare_layouts(ctx, oracle, op, dtype):
    la = dict(H.layouts(dtype, rot=0))
    lb = dict(H.layouts(dtype, rot=3))
    for na, nb in H.BINARY_LAYOUT_PAIRS:
        a, b = la[na], lb[nb]
        r = _run_both(lambda: oracle.compare(op, a, b),
                      lambda: getattr(B, op)(H.upload(ctx, a), H.upload(ctx, b)))
        if r[0] == "err":
            return
        H.assert_same("bool", H.download(r[2]), r[1].numpy(), what=f"{op}/{dtype}/{na},{nb}")


@pytest.mark.parametrize("dtype", ALL)
def test_where_layouts(ctx, oracle, dtype):
    la = dict(H.layouts(dtype, rot=0))
    lb = dict(H.layouts(dtype, rot=7))
    lc = dict(H.layouts("bool", rot=2))
    for nc, na, nb in [("contig", "contig", "contig"), ("transpose", "transpose", "transpose"),
                       ("broadcast", "flip", "slice"), ("scalar", "scalar", "scalar"),
                       ("empty", "empty", "empty"), ("permute3", "permute3", "permute3")]:
        c, a, b = lc[nc], la[na], lb[nb]
        want = oracle.where(c, a, b).numpy()
        got = H.download(B.where(H.upload(ctx, c), H.upload(ctx, a), H.upload(ctx, b)))
        assert np.array_equal(H.raw(got), H.raw(want)), f"where/{dtype}/{nc}"


def _cast_inputs(src):
    """Pool + the contract's cast edge cases: truncation, saturation, NaN, +-inf."""
    if src in H.FLOATS:
        extra = [1.9, -1.9, 0.5, -0.5, 300.7, -300.7, 70000.0, -70000.0, 3e9, -3e9, 1e19, -1e19, 1e30, -1e30,
                 float("nan"), float("inf"), float("-inf"), 127.5, 128.0, 255.9, 256.0, 2147483648.0, -0.0]
        if src in ("f8e4m3", "f8e5m2"):
            extra = [1.75, -1.75, 0.5, 240.0, -240.0, float("nan"), 0.0, -0.0]
        vals = np.concatenate([H.storage_to_float(src, H.pool(src)), np.array(extra)])
        return H.to_storage(src, vals)
    if src in H.COMPLEX:
        p = H.pool(src)
        extra = np.array([complex(1e30, 1), complex(float("nan"), 0), complex(-3.7, 9), complex(0, 0),
                          complex(0, 2)], dtype=p.dtype)
        return np.concatenate([p, extra])
    return H.pool(src)


@pytest.mark.parametrize("dst", ALL)
@pytest.mark.parametrize("src", ALL)
def test_cast_matrix(ctx, oracle, src, dst):
    data = _cast_inputs(src)
    n = (data.size // 2) * 2
    base = H.HostView(data[:n].copy(), src, [n // 2, 2])
    for name, hv in [("contig", base), ("transpose", base.permute([1, 0])),
                     ("flat", H.HostView(data.copy(), src, [data.size]))]:
        want = oracle.cast(hv, dst).numpy()
        got = H.download(B.cast(H.upload(ctx, hv), dst))
        # every conversion is a deterministic function of the stored bits: exact
        H.assert_same(dst, got, want, ulp=0, what=f"cast {src}->{dst} {name}")


@pytest.mark.parametrize("dtype", ALL)
def test_copy_contiguous_assign(ctx, oracle, dtype):
    for name, hv in H.layouts(dtype):
        t = H.upload(ctx, hv)
        got = B.copy(t)
        assert B.is_c_contiguous(got)
        assert np.array_equal(H.raw(H.download(got)), H.raw(hv.numpy())), f"copy/{dtype}/{name}"
        c = B.contiguous(t)
        assert B.is_c_contiguous(c) and c.offset == 0
    # assign: strided source into a strided destination sharing a base buffer
    lay = dict(H.layouts(dtype))
    src = lay["transpose"]
    dst_base = H.HostView(H.pool(dtype, 18).copy(), dtype, [3, 6])
    dst = dst_base.shrink([(0, 3), (1, 5)]).flip([0])
    dbase_t = H.upload(ctx, dst_base)
    dst_t = B.flip(B.shrink(dbase_t, [(0, 3), (1, 5)]), [True, False])
    B.assign(dst_t, H.upload(ctx, src))
    oracle.assign(dst, src)  # writes into dst_base.storage
    assert np.array_equal(H.raw(H.download(dbase_t)), H.raw(dst_base.numpy())), f"assign/{dtype}"


@pytest.mark.parametrize("dtype", ALL)
def test_full_and_scalar_operand(ctx, oracle, dtype):
    """`full` is a device-side fill; a rank-0 tensor expanded to stride 0 is how the
    frontend passes scalars (frontend.ml:360-361, 426-447)."""
    from raven_b200 import dtype as D
    p = H.pool(dtype, 18)
    val = p[3]
    t = B.full(ctx, D.of(dtype), [5, 7], val)
    got = H.download(t)
    assert got.shape == (5, 7) and (H.raw(got) == H.raw(np.full((5, 7), val))).all()
    if dtype == "bool":
        return
    a = H.HostView(np.tile(p, 4)[:60].copy(), dtype, [5, 12])
    s = H.HostView(p[3:4].copy(), dtype, []).expand([5, 12])
    want = oracle.binary("mul", a, s).numpy()
    gs = B.expand(B.full(ctx, D.of(dtype), [], val), [5, 12])
    got = H.download(B.mul(H.upload(ctx, a), gs))
    H.assert_same(dtype, got, want, ulp=(8 if dtype in H.COMPLEX else 0), what=f"mul_s/{dtype}")


# ---- large seeded arrays: the vector paths, tails, misalignment ------------------------------
def _rand(dtype, n, rng, lo=-4.0, hi=4.0):
    if dtype in H.FLOATS:
        return H.to_storage(dtype, rng.uniform(lo, hi, n))
    if dtype in H.COMPLEX:
        return (rng.uniform(lo, hi, n) + 1j * rng.uniform(lo, hi, n)).astype(H.np_storage(dtype))
    if dtype == "bool":
        return rng.integers(0, 2, n).astype(np.uint8)
    info = np.iinfo(H.np_storage(dtype))
    return rng.integers(info.min, info.max, n, dtype=H.np_storage(dtype), endpoint=True)


@pytest.mark.parametrize("dtype", ["f32", "f64", "f16", "bf16", "i32", "u8", "i64", "c32"])
@pytest.mark.parametrize("n", [1, 255, 4096 + 3, (1 << 20) + 17])
def test_large_binary_flat_and_offset(ctx, oracle, dtype, n):
    rng = np.random.default_rng(n)
    a = H.HostView(_rand(dtype, n + 1, rng), dtype, [n + 1])
    b = H.HostView(_rand(dtype, n + 1, rng), dtype, [n + 1])
    for off in (0, 1):  # offset 1 breaks 16-byte alignment: exercises the non-vector path
        av, bv = a.shrink([(off, n + off)]), b.shrink([(off, n + off)])
        for op in ("add", "mul"):
            want = oracle.binary(op, av, bv).numpy()
            got = H.download(getattr(B, op)(H.upload(ctx, av), H.upload(ctx, bv)))
            H.assert_same(dtype, got, want, ulp=(8 if dtype in H.COMPLEX else 0), what=f"{op}/{dtype}/n={n}/off={off}")


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("op", ["sin", "cos", "tan", "exp", "log", "sqrt", "asin", "acos", "atan", "sinh", "cosh",
                                "tanh", "erf", "recip"])
def test_large_unary_ulp(ctx, oracle, op, dtype):
    rng = np.random.default_rng(7)
    n = 1 << 20
    if op in ("asin", "acos"):
        x = rng.uniform(-1, 1, n)
    elif op in ("log", "sqrt"):
        x = np.exp(rng.uniform(-30, 30, n))
    elif op in ("exp", "sinh", "cosh"):
        x = rng.uniform(-30, 30, n)
    else:
        x = np.concatenate([rng.uniform(-4, 4, n // 2), rng.uniform(-1e4, 1e4, n // 2)])
    hv = H.HostView(H.to_storage(dtype, x), dtype, [n])
    want = oracle.unary(op, hv).numpy()
    got = H.download(getattr(B, op)(H.upload(ctx, hv)))
    # north_star's bound, no exception: f64 tanh follows glibc's own algorithm (nxc_tanh64.cuh)
    ulp = 0 if op in ("sqrt", "recip") else 2
    H.assert_same(dtype, got, want, ulp=ulp, what=f"{op}/{dtype}")


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("op", ["pow", "atan2", "fdiv", "mod"])
def test_large_binary_ulp(ctx, oracle, op, dtype):
    rng = np.random.default_rng(11)
    n = 1 << 18
    a = rng.uniform(0.01, 8, n) if op == "pow" else rng.uniform(-8, 8, n)
    b = rng.uniform(-8, 8, n)
    ha = H.HostView(H.to_storage(dtype, a), dtype, [n])
    hb = H.HostView(H.to_storage(dtype, b), dtype, [n])
    want = oracle.binary(op, ha, hb).numpy()
    got = H.download(getattr(B, BFN.get(op, op))(H.upload(ctx, ha), H.upload(ctx, hb)))
    H.assert_same(dtype, got, want, ulp=(0 if op in ("fdiv", "mod") else 2), what=f"{op}/{dtype}")


@pytest.mark.parametrize("dtype", ["f32", "i32", "f64", "u8"])
def test_large_2d_views(ctx, oracle, dtype):
    """Row / column broadcast, transposed operand, flipped and sliced rows at sizes
    that exercise the inner-vectorised strided kernel."""
    rng = np.random.default_rng(3)
    R, Cc = 257, 512
    a = H.HostView(_rand(dtype, R * Cc, rng), dtype, [R, Cc])
    row = H.HostView(_rand(dtype, Cc, rng), dtype, [1, Cc]).expand([R, Cc])
    col = H.HostView(_rand(dtype, R, rng), dtype, [R, 1]).expand([R, Cc])
    at = H.HostView(_rand(dtype, R * Cc, rng), dtype, [Cc, R]).permute([1, 0])
    sl = H.HostView(_rand(dtype, R * (Cc + 8), rng), dtype, [R, Cc + 8]).shrink([(0, R), (4, Cc + 4)])
    fl = a.flip([0, 1])
    for name, other in [("row", row), ("col", col), ("transposed", at), ("slice", sl), ("flip", fl)]:
        want = oracle.binary("add", a, other).numpy()
        got = H.download(B.add(H.upload(ctx, a), H.upload(ctx, other)))
        H.assert_same(dtype, got, want, ulp=0, what=f"add/{dtype}/{name}")


def test_rank32(ctx, oracle):
    """Rank-32 add (backend_contract.ml:3012-3042)."""
    shape = [2, 2] + [1] * 30
    a = H.HostView(np.array([1, 2, 3, 4], dtype=np.float32), "f32", shape)
    b = H.HostView(np.array([10, 20, 30, 40], dtype=np.float32), "f32", shape)
    got = H.download(B.add(H.upload(ctx, a), H.upload(ctx, b)))
    assert got.reshape(-1).tolist() == [11, 22, 33, 44]
    perm = list(range(32))
    perm[0], perm[1] = 1, 0
    got = H.download(B.add(B.permute(H.upload(ctx, a), perm), H.upload(ctx, b)))
    assert got.reshape(-1).tolist() == [11, 23, 32, 44]


@pytest.mark.parametrize("dtype", ["bf16", "f16", "i8", "u8", "i16", "u16", "bool"])
def test_transposed_views_of_narrow_dtypes(ctx, oracle, dtype):
    """The 128 x 128 swizzled-tile kernel for 1- and 2-byte elements (nxc_map_tiledn_kernel): one or
    both operands transposed, extents that are and are not multiples of the tile (but of the 16-byte
    vector), a batch dim, a row-broadcast straight operand, contiguous(transpose); and extents that
    do not divide the vector, which must fall back to the other tiled kernels. Bit-exact."""
    rng = np.random.default_rng(31)
    ops = ("max",) if dtype == "bool" else ("add", "mul")
    for (R, Cc) in [(256, 384), (144, 208), (130, 100), (48, 1040)]:
        a = H.HostView(_rand(dtype, R * Cc, rng), dtype, [R, Cc])
        at = H.HostView(_rand(dtype, R * Cc, rng), dtype, [Cc, R]).permute([1, 0])
        bt = H.HostView(_rand(dtype, R * Cc, rng), dtype, [Cc, R]).permute([1, 0])
        row = H.HostView(_rand(dtype, Cc, rng), dtype, [1, Cc]).expand([R, Cc])
        for op in ops:
            for name, (x, y) in {"a + bT": (a, at), "aT + bT": (at, bt), "aT + row": (at, row)}.items():
                want = oracle.binary(op, x, y).numpy()
                got = H.download(getattr(B, op)(H.upload(ctx, x), H.upload(ctx, y)))
                H.assert_same(dtype, got, want, ulp=0, what=f"{op}/{dtype}/{R}x{Cc}/{name}")
        want = oracle.copy(at).numpy()
        got = H.download(B.contiguous(H.upload(ctx, at)))
        assert np.array_equal(H.raw(got), H.raw(want)), f"contiguous(transpose)/{dtype}/{R}x{Cc}"
    # batched: [3, R, C] + transposed-in-the-last-two-dims view of [3, C, R]
    R, Cc = 160, 272
    a3 = H.HostView(_rand(dtype, 3 * R * Cc, rng), dtype, [3, R, Cc])
    b3 = H.HostView(_rand(dtype, 3 * R * Cc, rng), dtype, [3, Cc, R]).permute([0, 2, 1])
    op = ops[0]
    want = oracle.binary(op, a3, b3).numpy()
    got = H.download(getattr(B, op)(H.upload(ctx, a3), H.upload(ctx, b3)))
    H.assert_same(dtype, got, want, ulp=0, what=f"{op}/{dtype}/batched transposed")
