"""packages/nx-cuda/lib/nx_cuda_stubs.c EXECUTED on the GPU (tests/stubs_harness.py builds OCaml
values by hand and a ~100-line runtime shim stands in for the OCaml runtime): every stub is driven
with the contract suite's layout matrix and compared with the oracle, including the exception
class and "<op>: <message>" text each failure raises, the record's slot order, the kind -> tag
table for all 19 dtypes and the custom-block lifetime (finalizer -> nxc_free). What remains
unverified about the OCaml package is only the OCaml text of nx_backend.ml.
Reference for what the stubs must do: packages/nx/lib/backend_c/test/test_backend_c.ml, nx_c.h:406-434."""
import numpy as np
import pytest

from raven_b200._lib import Failure, InvalidArgument
from tests import harness as H
from tests import stubs_harness as S

pytestmark = pytest.mark.gpu

ALL = list(H.FLOATS) + list(H.INTS) + list(H.COMPLEX) + ["bool"]


@pytest.fixture(scope="module")
def sb():
    S.build()
    return S.StubBackend()


def _both(ofn, gfn):
    try:
        want = ofn()
    except Exception as e:
        with pytest.raises(InvalidArgument if e.kind == "Invalid_argument" else Failure) as ei:
            gfn()
        assert str(ei.value).startswith(e.msg), (str(ei.value), e.msg)
        return None, None
    return want, gfn()


def _ulp(dtype, exact):
    if dtype in H.INTS or dtype == "bool" or exact:
        return 0
    return 2 if dtype in ("f32", "f64") else 1


@pytest.mark.parametrize("dtype", ALL)
def test_map_family_through_the_stubs(sb, oracle, dtype):
    la, lb = dict(H.layouts(dtype, rot=0)), dict(H.layouts(dtype, rot=5))
    for op in ("neg", "sqrt", "sin", "abs", "floor"):
        for name, hv in la.items():
            want, got = _both(lambda: oracle.unary(op, hv), lambda: sb.unary(op, sb.upload(hv)))
            if want is None:
                break
            if dtype in H.COMPLEX and op not in ("neg", "abs"):
                H.assert_close(dtype, sb.download(got), want.numpy(), rel=1e-5 if dtype == "c32" else 1e-11,
                               abs_=1e-5 if dtype == "c32" else 1e-11, what=f"{op}/{dtype}/{name}")
            else:
                H.assert_same(dtype, sb.download(got), want.numpy(), ulp=_ulp(dtype, op in ("neg", "sqrt", "abs", "floor")),
                              what=f"{op}/{dtype}/{name}")
    for op in ("add", "mul", "idiv", "max", "xor", "shl"):
        for na, nb in H.BINARY_LAYOUT_PAIRS:
            want, got = _both(lambda: oracle.binary(op, la[na], lb[nb]),
                              lambda: sb.binary(op, sb.upload(la[na]), sb.upload(lb[nb])))
            if want is None:
                break
            if dtype in H.COMPLEX and op == "mul":
                H.assert_close(dtype, sb.download(got), want.numpy(), rel=1e-5 if dtype == "c32" else 1e-11, abs_=1e-11,
                               what=f"{op}/{dtype}")
            else:
                H.assert_same(dtype, sb.download(got), want.numpy(), ulp=0, what=f"{op}/{dtype}/{na},{nb}")
    for na, nb in H.BINARY_LAYOUT_PAIRS:
        want, got = _both(lambda: oracle.compare("cmplt", la[na], lb[nb]),
                          lambda: sb.compare("cmplt", sb.upload(la[na]), sb.upload(lb[nb])))
        if want is None:
            break
        H.assert_same("bool", sb.download(got), want.numpy(), what=f"cmplt/{dtype}/{na},{nb}")
    lc = dict(H.layouts("bool", rot=2))
    for nc, na, nb in (("contig", "contig", "contig"), ("broadcast", "flip", "slice"), ("permute3", "permute3", "permute3")):
        want = oracle.where(lc[nc], la[na], lb[nb]).numpy()
        got = sb.download(sb.where(sb.upload(lc[nc]), sb.upload(la[na]), sb.upload(lb[nb])))
        assert np.array_equal(H.raw(got), H.raw(want)), f"where/{dtype}/{nc}"
    # copy into a strided destination that shares its base (assign), as the contract does
    base = H.HostView(np.zeros(18, dtype=H.np_storage(dtype)), dtype, [3, 6])
    dst = base.shrink([(0, 3), (1, 5)]).flip([0])
    tb = sb.upload(base)
    src = la["transpose"]                                   # [3, 4] through a transposed view
    sb.assign(S.ST(tb.buf, dst.shape, dst.strides, dst.offset, dtype, tb.ctx, tb.elems), sb.upload(src))
    oracle.assign(dst, src)
    assert np.array_equal(H.raw(sb.to_host(tb)), H.raw(base.storage)), f"assign/{dtype}"


@pytest.mark.parametrize("src", ALL)
def test_cast_through_the_stubs_every_tag(sb, oracle, src):
    """dst runs over all dtypes: exercises slot 4 (the Dtype.Packed tag) of the record for each."""
    hv = dict(H.layouts(src))["transpose"]
    for dst in ALL:
        want = oracle.cast(hv, dst).numpy()
        got = sb.download(sb.cast(sb.upload(hv), dst))
        assert np.array_equal(H.raw(got), H.raw(want)), f"cast {src}->{dst}"


@pytest.mark.parametrize("dtype", ["f32", "f64", "i32", "u8", "bf16", "i64", "c32", "bool"])
def test_fold_family_through_the_stubs(sb, oracle, dtype):
    la = dict(H.layouts(dtype, include_degenerate=False))
    for name in ("contig", "transpose", "slice", "flip", "permute3"):
        hv = la[name]
        for op in ("sum", "prod", "max", "min"):
            for axes in ([0], [len(hv.shape) - 1], list(range(len(hv.shape)))):
                want, got = _both(lambda: oracle.reduce(op, hv, axes), lambda: sb.reduce(op, sb.upload(hv), axes))
                if want is None:
                    continue
                if dtype in H.INTS or dtype == "bool" or op in ("max", "min"):
                    H.assert_same(dtype, sb.download(got), want.numpy(), what=f"reduce {op}/{dtype}/{name}/{axes}")
                else:
                    H.assert_close(dtype, sb.download(got), want.numpy(), rel=2e-2 if dtype == "bf16" else 1e-5, abs_=1e-5,
                                   what=f"reduce {op}/{dtype}/{name}/{axes}")
        for op in ("argmax", "argmin"):
            for keep in (False, True):
                want, got = _both(lambda: oracle.argreduce(op, hv, 0, keep), lambda: sb.argreduce(op, sb.upload(hv), 0, keep))
                if want is not None:
                    H.assert_same("i32", sb.download(got), want.numpy(), what=f"{op}/{dtype}/{name}")
        want, got = _both(lambda: oracle.scan("sum", hv, 0), lambda: sb.scan("sum", sb.upload(hv), 0))
        if want is not None and (dtype in H.INTS or dtype in ("f32", "f64")):
            H.assert_same(dtype, sb.download(got), want.numpy(), what=f"cumsum/{dtype}/{name}")


@pytest.mark.parametrize("dtype", ["f32", "f64", "i32", "bf16", "c32"])
def test_matmul_through_the_stubs(sb, oracle, dtype):
    rng = np.random.default_rng(3)
    m, k, n = 37, 19, 23
    A = H.to_storage(dtype, rng.integers(-4, 5, (m, k)).astype(np.float64))
    Bm = H.to_storage(dtype, rng.integers(-4, 5, (k, n)).astype(np.float64))
    a = H.HostView(A.reshape(-1).copy(), dtype, [m, k])
    b = H.HostView(Bm.reshape(-1).copy(), dtype, [k, n])
    at = H.HostView(np.ascontiguousarray(A.T).reshape(-1), dtype, [k, m]).permute([1, 0])
    for x, y in ((a, b), (at, b)):
        want = oracle.matmul(x, y).numpy()
        got = sb.download(sb.matmul(sb.upload(x), sb.upload(y)))
        assert np.array_equal(H.raw(got), H.raw(want)), f"matmul/{dtype} (small integers: exact in every dtype)"
    a3 = H.HostView(np.tile(A.reshape(-1), 2), dtype, [2, m, k])
    want = oracle.matmul(a3, b).numpy()
    assert np.array_equal(H.raw(sb.download(sb.matmul(sb.upload(a3), sb.upload(b)))), H.raw(want)), "batch broadcast"
    with pytest.raises(InvalidArgument, match="^matmul: shape mismatch"):
        sb.matmul(sb.upload(a), sb.upload(a))
    f = H.HostView(np.zeros(4, dtype=np.float32), "f32", [2, 2])
    d = H.HostView(np.zeros(4, dtype=np.float64), "f64", [2, 2])
    with pytest.raises(Failure, match="^matmul: matmul operands must share one dtype"):
        x, y = sb.upload(f), sb.upload(d)
        out = sb.create("f32", [2, 2])
        sb._call("nx_cuda_matmul", 3, lambda ar: [out.record(ar), x.record(ar), y.record(ar)])


def test_move_family_through_the_stubs(sb, oracle):
    rng = np.random.default_rng(5)
    x = H.HostView(rng.integers(-50, 50, 24).astype(np.int32), "i32", [4, 6])
    xt = x.permute([1, 0])
    fill = np.array([7], dtype=np.int32)
    want = oracle.pad(xt, [(1, 2), (0, 3)], H.HostView(fill, "i32", [])).numpy()
    assert np.array_equal(sb.download(sb.pad(sb.upload(xt), [(1, 2), (0, 3)], fill)), want)
    want = oracle.cat([x, x.flip([0]), x], 0).numpy()
    assert np.array_equal(sb.download(sb.cat([sb.upload(x), sb.upload(x.flip([0])), sb.upload(x)], 0)), want)
    idx = H.HostView(rng.integers(-4, 4, 18).astype(np.int32), "i32", [3, 6])
    want = oracle.gather(x, idx, 0).numpy()
    assert np.array_equal(sb.download(sb.gather(sb.upload(x), sb.upload(idx), 0)), want)
    bad = H.HostView(np.full(18, 9, dtype=np.int32), "i32", [3, 6])
    with pytest.raises(Failure, match="^gather: index out of bounds"):
        sb.gather(sb.upload(x), sb.upload(bad), 0)
    upd = H.HostView(rng.integers(-9, 9, 18).astype(np.int32), "i32", [3, 6])
    for mode in ("set", "add"):
        want = oracle.scatter(x, idx, upd, 0, mode).numpy()
        assert np.array_equal(sb.download(sb.scatter(sb.upload(x), sb.upload(idx), sb.upload(upd), 0, mode)), want), mode
    xf = H.HostView(rng.standard_normal(40).astype(np.float32), "f32", [5, 8])
    xf.storage[3] = np.nan
    for desc in (False, True):
        assert np.array_equal(H.raw(sb.download(sb.sort(sb.upload(xf), 1, desc))), H.raw(oracle.sort(xf, 1, desc).numpy()))
        assert np.array_equal(sb.download(sb.sort(sb.upload(xf), 1, desc, arg=True)), oracle.argsort(xf, 1, desc).numpy())
    key = H.HostView(np.array([0x13198A2E, 0x03707344], dtype=np.uint32).view(np.int32), "i32", [1, 2])
    ctr = H.HostView(np.array([0x243F6A88, 0x85A308D3], dtype=np.uint32).view(np.int32), "i32", [1, 2])
    got = sb.download(sb.threefry(sb.upload(key), sb.upload(ctr))).view(np.uint32)
    assert got.tolist() == [[0xC4923A9C, 0x483DF7A0]]          # Random123 KAT (backend_contract.ml:1663-1666)
    img = H.HostView(rng.standard_normal(2 * 6 * 7).astype(np.float32), "f32", [2, 6, 7])
    want = oracle.unfold(img, [3, 2], [1, 2], [1, 1], [(1, 1), (0, 1)]).numpy()
    for bc in (False, True):   # native and bytecode entry points (> 5 arguments)
        got = sb.download(sb.unfold(sb.upload(img), [3, 2], [1, 2], [1, 1], [(1, 1), (0, 1)], bytecode=bc))
        assert np.array_equal(H.raw(got), H.raw(want)), f"unfold bytecode={bc}"


def test_full_transfer_errors_and_lifetime(sb, oracle):
    for dtype, val in (("f32", 2.5), ("bf16", -1.5), ("i64", -7), ("u8", 200), ("bool", 1), ("c64", 1.0 - 2.0j)):
        st = H.to_storage(dtype, np.array([val]))
        got = sb.download(sb.full(dtype, [3, 5], st))
        assert got.shape == (3, 5) and (H.raw(got).reshape(15, -1) == H.raw(st)).all(), dtype
    # unknown op code: the engine's status, the stub must not index past its name table
    x = sb.upload(H.HostView(np.ones(4, dtype=np.float32), "f32", [4]))
    with pytest.raises(Failure, match="^op: unknown operation code"):
        sb.raw_map1(99, x)
    # dtype not supported for the op: same class and text as the reference funnel
    b = sb.upload(H.HostView(np.ones(4, dtype=np.uint8), "bool", [4]))
    with pytest.raises(Failure, match="^sin: dtype not supported for this operation"):
        sb.unary("sin", b)
    with pytest.raises(InvalidArgument, match="^reduce_sum: reduce axes must be strictly increasing and in range"):
        out = sb.create("f32", [])
        sb._call("nx_cuda_reduce", 4, lambda a: [S.val_int(0), out.record(a), x.record(a), a.ints([3])])
    # linalg: numeric failure text that the veneer lifts to Linalg_error
    spd = np.array([[4, 2], [2, 3]], dtype=np.float32)
    got = sb.download(sb.cholesky(sb.upload(H.HostView(spd.reshape(-1).copy(), "f32", [2, 2]))))
    assert np.allclose(got @ got.T, spd, atol=1e-6)
    with pytest.raises(Failure, match="^cholesky: matrix is not positive definite"):
        sb.cholesky(sb.upload(H.HostView(np.array([1, 2, 2, 1], dtype=np.float32), "f32", [2, 2])))
    z = H.HostView((np.arange(16) + 1j * np.arange(16)[::-1]).astype(np.complex64), "c32", [16])
    H.assert_close("c32", sb.download(sb.fft(sb.upload(z), [0])), oracle.fft(z, [0]).numpy(), rel=2e-6, abs_=2e-5, what="fft")
    # custom-block lifetime: dropping the wrappers runs the finalizer (nxc_free) and frees the block
    import gc
    del got, x, b, out
    gc.collect()                       # wrappers of the tensors above go first
    before = S.lib().nxstub_live()
    ts = [sb.create("f32", [1 << 20]) for _ in range(8)]
    assert S.lib().nxstub_live() == before + 8
    assert S.lib().nxstub_custom_identifier(ts[0].buf.value) == b"nx_cuda.devbuf"
    del ts
    gc.collect()
    assert S.lib().nxstub_live() == before


def test_capture_stubs_replay_a_step(sb, oracle):
    """nx_cuda_capture_begin / _end / nx_cuda_graph_launch: ops issued through the stubs between begin
    and end are recorded, the graph value is a custom block (finalizer -> nxc_graph_destroy), a replay
    rewrites the outputs in place, and a blocking call inside a capture raises Failure."""
    rng = np.random.default_rng(8)
    a = H.HostView(rng.uniform(-2, 2, 4096).astype(np.float32), "f32", [64, 64])
    b = H.HostView(rng.uniform(-2, 2, 4096).astype(np.float32), "f32", [64, 64])
    ta, tb = sb.upload(a), sb.upload(b)
    sb.capture_begin()
    r = sb.binary("mul", sb.binary("add", ta, tb), sb.permute(ta, [1, 0]))
    s = sb.reduce("sum", r, [1])
    g = sb.capture_end()
    assert S.lib().nxstub_custom_identifier(g.value) == b"nx_cuda.graph"
    sb.graph_launch(g)
    want_r = oracle.binary("mul", oracle.binary("add", a, b), a.permute([1, 0]))
    H.assert_same("f32", sb.download(r), want_r.numpy(), ulp=0, what="captured mul(add)")
    H.assert_close("f32", sb.download(s), oracle.reduce("sum", want_r, [1]).numpy(), rel=1e-5, abs_=1e-5, what="captured sum")
    # refreshed input, in place; the replay sees it
    a2 = H.HostView(rng.uniform(-2, 2, 4096).astype(np.float32), "f32", [64, 64])
    sb.assign(ta, sb.upload(a2))
    sb.graph_launch(g)
    want_r = oracle.binary("mul", oracle.binary("add", a2, b), a2.permute([1, 0]))
    H.assert_same("f32", sb.download(r), want_r.numpy(), ulp=0, what="replay over a refreshed input")
    sb.capture_begin()
    with pytest.raises(Failure, match="^to_host: operation not allowed while a step is being captured"):
        sb.to_host(ta)
    sb.capture_end()
    sb.sync()
