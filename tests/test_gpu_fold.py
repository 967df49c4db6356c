"""GPU parity of the fold family (reduce sum/prod/max/min, argmax/argmin) against
the oracle through the C ABI.

Tolerances (north_star): integer/bool reductions and every argmax/argmin are
bit-exact; float max/min are exact (NaN positions must agree); float sum/prod are
within 1e-5 relative for f32 (1e-11 f64, one storage ulp-ish for f16/bf16, scaled
by the reduced count as the contract does, backend_contract.ml:907-913).
"""
import numpy as np
import pytest

import raven_b200.backend as B
from raven_b200 import Failure, InvalidArgument
from tests import harness as H
from tests.test_gpu_map import _rand

pytestmark = pytest.mark.gpu

ORD = list(H.FLOATS) + list(H.INTS) + ["bool"]
ARITH = list(H.FLOATS) + list(H.INTS) + list(H.COMPLEX)
REL = {"f64": 1e-11, "f32": 1e-5, "f16": 5e-3, "bf16": 3e-2, "f8e4m3": 0.13, "f8e5m2": 0.26, "c32": 1e-5,
       "c64": 1e-11}


def _check_reduce(ctx, oracle, op, dtype, hv, axes, what):
    try:
        want = oracle.reduce(op, hv, axes)
    except Exception as e:
        with pytest.raises(InvalidArgument if e.kind == "Invalid_argument" else Failure) as ei:
            B.reduce(H.upload(ctx, hv), op, axes)
        assert str(ei.value).startswith(e.msg.split(":")[0]), (str(ei.value), e.msg)
        return
    got = H.download(B.reduce(H.upload(ctx, hv), op, axes))
    want = want.numpy()
    if dtype in H.INTS or dtype == "bool" or op in ("max", "min"):
        H.assert_same(dtype, got, want, ulp=0, what=what)
    else:
        n = 1
        for a in axes:
            n *= hv.shape[a]
        rel = REL[dtype] * max(1.0, np.sqrt(n))
        H.assert_close(dtype, got, want, rel=rel, abs_=rel, what=what)


@pytest.mark.parametrize("dtype", ORD + list(H.COMPLEX))
@pytest.mark.parametrize("op", ["sum", "prod", "max", "min"])
def test_reduce_layouts(ctx, oracle, op, dtype):
    for name, hv in H.layouts(dtype):
        nd = len(hv.shape)
        opts = [[]] if nd == 0 else [[0], [nd - 1], list(range(nd))]
        for axes in opts:
            _check_reduce(ctx, oracle, op, dtype, hv, axes, f"{op}/{dtype}/{name}/{axes}")


@pytest.mark.parametrize("dtype", ORD)
@pytest.mark.parametrize("op", ["argmax", "argmin"])
def test_argreduce_layouts(ctx, oracle, op, dtype):
    for name, hv in H.layouts(dtype, include_degenerate=False):
        nd = len(hv.shape)
        for axis in (0, nd - 1):
            for keep in (False, True):
                want = oracle.argreduce(op, hv, axis, keep).numpy()
                got = H.download(getattr(B, op)(H.upload(ctx, hv), axis, keep))
                H.assert_same("i32", got, want, what=f"{op}/{dtype}/{name}/axis{axis}/keep{keep}")


def test_empty_axis_errors(ctx):
    from raven_b200 import dtype as D
    e = B.buffer(ctx, D.float32, [0, 4])
    for op in ("max", "min"):
        with pytest.raises(InvalidArgument, match=f"reduce_{op}: reduction over an empty axis has no identity"):
            B.reduce(e, op, [0])
    with pytest.raises(InvalidArgument, match="argmax: argument reduction over an empty axis"):
        B.argmax(e, 0)
    # sum/prod over an empty extent store the identity (nx_c.h:489-495)
    assert H.download(B.reduce(e, "sum", [0])).tolist() == [0, 0, 0, 0]
    assert H.download(B.reduce(e, "prod", [0])).tolist() == [1, 1, 1, 1]
    # the driver re-checks what a binding could get wrong (nx_c_engine.c:1085-1093)
    import ctypes
    x = B.buffer(ctx, D.float32, [3, 4])
    out = B.buffer(ctx, D.float32, [])
    do, dx = out._desc(), x._desc()
    ax = (ctypes.c_int * 2)(1, 0)
    st = ctx._lib.nxc_reduce(ctx.ptr, 0, ctypes.byref(do), ctypes.byref(dx), ax, 2)
    assert st == b"reduce axes must be strictly increasing and in range"
    assert ctx._lib.nxc_status_is_invalid_argument(st) == 1


SHAPES = [  # (shape, axes)
    ([1 << 20], [0]),
    ([1024, 1024], [0]), ([1024, 1024], [1]), ([1024, 1024], [0, 1]),
    ([4096, 256], [0]), ([4096, 256], [1]), ([256, 4096], [0]), ([256, 4096], [1]),
    ([37, 1001], [0]), ([37, 1001], [1]), ([3, 5, 7, 11], [1, 3]), ([3, 5, 7, 11], [0, 2]),
    ([64, 3, 1000], [1]), ([2, 500000], [0]), ([500000, 2], [0]), ([500000, 2], [1]),
]


@pytest.mark.parametrize("dtype", ["f32", "f64", "i32", "u8", "bf16", "i64"])
@pytest.mark.parametrize("op", ["sum", "max", "min", "prod"])
def test_reduce_large(ctx, oracle, op, dtype):
    rng = np.random.default_rng(5)
    for shape, axes in SHAPES:
        n = int(np.prod(shape))
        if op == "prod" and dtype in H.FLOATS:
            data = H.to_storage(dtype, rng.uniform(0.9, 1.1, n))
        elif dtype in H.FLOATS:
            data = H.to_storage(dtype, rng.uniform(-1, 1, n))
        else:
            data = _rand(dtype, n, rng)
        hv = H.HostView(data, dtype, shape)
        _check_reduce(ctx, oracle, op, dtype, hv, axes, f"{op}/{dtype}/{shape}/{axes}")
    # transposed and sliced views of a matrix
    data = H.to_storage(dtype, rng.uniform(-1, 1, 600 * 700)) if dtype in H.FLOATS else _rand(dtype, 600 * 700, rng)
    base = H.HostView(data, dtype, [600, 700])
    for name, hv in [("T", base.permute([1, 0])), ("slice", base.shrink([(3, 590), (5, 690)])),
                     ("flip", base.flip([0, 1]))]:
        for axes in ([0], [1], [0, 1]):
            if op == "prod" and dtype in H.FLOATS:
                continue
            _check_reduce(ctx, oracle, op, dtype, hv, axes, f"{op}/{dtype}/{name}/{axes}")


@pytest.mark.parametrize("dtype", ["f32", "f64", "i32", "u32", "u8", "bf16"])
@pytest.mark.parametrize("op", ["argmax", "argmin"])
def test_argreduce_large(ctx, oracle, op, dtype):
    rng = np.random.default_rng(9)
    for shape, axis in [([1 << 20], 0), ([1024, 1024], 0), ([1024, 1024], 1), ([4096, 256], 1),
                        ([256, 4096], 0), ([37, 1001], 1), ([64, 3, 1000], 1), ([500000, 2], 0)]:
        n = int(np.prod(shape))
        if dtype in H.FLOATS:
            vals = np.round(rng.uniform(-50, 50, n))  # many exact ties
            data = H.to_storage(dtype, vals)
        else:
            data = _rand(dtype, n, rng)
        hv = H.HostView(data, dtype, shape)
        want = oracle.argreduce(op, hv, axis).numpy()
        got = H.download(getattr(B, op)(H.upload(ctx, hv), axis))
        H.assert_same("i32", got, want, what=f"{op}/{dtype}/{shape}/{axis}")
    # transposed view
    hv = H.HostView(_rand(dtype, 300 * 400, rng) if dtype not in H.FLOATS else
                    H.to_storage(dtype, np.round(rng.uniform(-9, 9, 300 * 400))), dtype, [300, 400]).permute([1, 0])
    for axis in (0, 1):
        want = oracle.argreduce(op, hv, axis).numpy()
        got = H.download(getattr(B, op)(H.upload(ctx, hv), axis))
        H.assert_same("i32", got, want, what=f"{op}/{dtype}/T/{axis}")


def test_nan_semantics(ctx, oracle):
    """NaN sticks for max/min; the FIRST NaN wins argmax/argmin; ties keep the first
    index (backend_contract.ml regression group, nx_c_fold.c:80-101)."""
    x = np.arange(4096, dtype=np.float32)
    x[100] = np.nan
    x[3000] = np.nan
    x[5] = 1e9
    x[6] = 1e9
    hv = H.HostView(x, "f32", [4096])
    t = H.upload(ctx, hv)
    assert np.isnan(H.download(B.reduce(t, "max", [0])))
    assert np.isnan(H.download(B.reduce(t, "min", [0])))
    assert int(H.download(B.argmax(t, 0))) == 100 == int(oracle.argreduce("argmax", hv, 0).numpy())
    assert int(H.download(B.argmin(t, 0))) == 100
    y = x.copy()
    y[[100, 3000]] = 0
    assert int(H.download(B.argmax(H.upload(ctx, H.HostView(y, "f32", [4096])), 0))) == 5


def test_accumulator_witnesses(ctx):
    """2^25 ones sum exactly in f32; f16 4096 ones and bf16 512 ones do not stall at
    the storage type's integer ceiling (backend_contract.ml:2611-2644)."""
    from raven_b200 import dtype as D
    ones = B.full(ctx, D.float32, [1 << 25], 1.0)
    assert float(H.download(B.reduce(ones, "sum", [0]))) == float(1 << 25)
    h = B.full(ctx, D.float16, [4096], 1.0)
    assert H.download(B.reduce(h, "sum", [0])).view(np.float16) == np.float16(4096)
    b = B.full(ctx, D.bfloat16, [512], 1.0)
    assert int(H.download(B.reduce(b, "sum", [0]))) == 0x4400  # 512.0 in bf16
    # bool: max = any, min = all (backend_contract.ml:2425-2432)
    m = np.ones(1000, dtype=np.uint8)
    m[777] = 0
    t = H.upload(ctx, H.HostView(m, "bool", [1000]))
    assert int(H.download(B.reduce(t, "min", [0]))) == 0 and int(H.download(B.reduce(t, "max", [0]))) == 1
    # unsigned argmax over a high-bit u32 (backend_contract.ml:2773-2817)
    u = np.array([3, 0x80000000, 7, 0xFFFFFFF0, 1], dtype=np.uint32)
    assert int(H.download(B.argmax(H.upload(ctx, H.HostView(u, "u32", [5])), 0))) == 3


def test_full_size_properties(ctx):
    """BASELINE.json's 2^28-element size through size-independent properties: a sum
    of ones, linearity sum(a+a) == 2*sum(a), argmax of a planted maximum."""
    from raven_b200 import dtype as D
    n = 1 << 28
    ones = B.full(ctx, D.float32, [n], 1.0)
    assert float(H.download(B.reduce(ones, "sum", [0]))) == float(n)
    two = B.add(ones, ones)
    assert float(H.download(B.reduce(two, "sum", [0]))) == float(2 * n)
    s2 = H.download(B.reduce(B.reshape(two, [1 << 14, 1 << 14]), "sum", [0]))
    assert (s2 == float(2 << 14)).all()
    s3 = H.download(B.reduce(B.reshape(two, [1 << 14, 1 << 14]), "sum", [1]))
    assert (s3 == float(2 << 14)).all()
    # plant a maximum via assign into a 1-element slice
    pos = 123456789
    B.assign(B.shrink(ones, [(pos, pos + 1)]), B.full(ctx, D.float32, [1], 2.0))
    assert int(H.download(B.argmax(ones, 0))) == pos
    assert float(H.download(B.reduce(ones, "max", [0]))) == 2.0
    i = B.full(ctx, D.int32, [n], 3)
    assert int(H.download(B.reduce(i, "sum", [0]))) == np.int32((3 * n) & 0xFFFFFFFF)
