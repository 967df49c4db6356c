"""GPU parity of the "next" rows of the scope table:
pad, cat, gather, scatter (Set last-wins / Add), threefry, cumulative scans --
against the oracle through the C ABI. Everything here is bit-exact except float
scatter-Add with duplicate targets (atomic order; 1e-5 relative) -- scans are
sequential per slice on both sides, so even float scans are bit-identical.
"""
import numpy as np
import pytest

import raven_b200.backend as B
from raven_b200 import Failure, InvalidArgument
from raven_b200 import dtype as DT
from tests import harness as H
from tests.test_gpu_map import _rand

pytestmark = pytest.mark.gpu

DTS = ["f32", "f64", "i32", "u8", "bf16", "f16", "c32", "bool", "i64", "u16"]


def _views(dt):
    return {"contig": H.HostView(H.pool(dt, 18).copy(), dt, [3, 6]),
            "T": H.HostView(H.pool(dt, 18).copy(), dt, [6, 3]).permute([1, 0]),
            "slice": H.HostView(np.tile(H.pool(dt, 18), 2).copy(), dt, [3, 12]).shrink([(0, 3), (2, 8)])}


@pytest.mark.parametrize("dt", DTS)
def test_pad_cat(ctx, oracle, dt):
    vs = _views(dt)
    fill_hv = H.HostView(H.pool(dt, 18)[4:5].copy(), dt, [])
    fill_val = H.pool(dt, 18)[4]
    for name, v in vs.items():
        for padding in ([(1, 2), (0, 3)], [(0, 0), (2, 0)], [(0, 0), (0, 0)]):
            want = oracle.pad(v, padding, fill_hv).numpy()
            got = H.download(B.pad(H.upload(ctx, v), padding, fill_val))
            assert np.array_equal(H.raw(got), H.raw(want)), f"pad/{dt}/{name}/{padding}"
        want = oracle.cat([v, vs["contig"], v], 0).numpy()
        got = H.download(B.cat([H.upload(ctx, v), H.upload(ctx, vs["contig"]), H.upload(ctx, v)], 0))
        assert np.array_equal(H.raw(got), H.raw(want)), f"cat0/{dt}/{name}"
        want = oracle.cat([v, vs["T"]], 1).numpy()
        got = H.download(B.cat([H.upload(ctx, v), H.upload(ctx, vs["T"])], -1))
        assert np.array_equal(H.raw(got), H.raw(want)), f"cat1/{dt}/{name}"


@pytest.mark.parametrize("dt", DTS)
def test_gather_scatter(ctx, oracle, dt):
    rng = np.random.default_rng(4)
    for name, v in _views(dt).items():
        for axis, hi in ((0, 3), (1, 6)):
            shape = [4, 6] if axis == 0 else [3, 5]
            idx = H.HostView(rng.integers(-hi, hi, shape[0] * shape[1]).astype(np.int32), "i32", shape)
            want = oracle.gather(v, idx, axis).numpy()
            got = H.download(B.gather(H.upload(ctx, v), H.upload(ctx, idx), axis))
            assert np.array_equal(H.raw(got), H.raw(want)), f"gather/{dt}/{name}/{axis}"
            upd = H.HostView(np.resize(H.pool(dt, 18), shape[0] * shape[1]).copy(), dt, shape)
            for mode in ("set", "add"):
                want = oracle.scatter(v, idx, upd, axis, mode).numpy()
                got = H.download(B.scatter(H.upload(ctx, v), H.upload(ctx, idx), H.upload(ctx, upd), axis, mode))
                if mode == "add" and dt in H.FLOATS + H.COMPLEX:
                    H.assert_close(dt, got, want, rel=3e-2 if dt in ("bf16", "f16") else 1e-5, abs_=1e-5,
                                   what=f"scatter-add/{dt}/{name}/{axis}")
                else:
                    assert np.array_equal(H.raw(got), H.raw(want)), f"scatter-{mode}/{dt}/{name}/{axis}"
    base = _views(dt)["contig"]
    bad = H.HostView(np.array([0, 9, 1] * 6, dtype=np.int32), "i32", [3, 6])
    with pytest.raises(Failure, match="gather: index out of bounds for the gathered/scattered axis"):
        B.gather(H.upload(ctx, base), H.upload(ctx, bad), 0)


def test_embedding_shaped_gather_and_scatter_add(ctx, oracle):
    """take -> gather with a column-broadcast index, backward = scatter Add into zeros
    (kaun/lib/embedding.ml, rune/lib/reverse.ml:521-526)."""
    rng = np.random.default_rng(0)
    V, Dm, T = 1000, 64, 4096
    table = H.HostView(rng.standard_normal(V * Dm).astype(np.float32), "f32", [V, Dm])
    ids = H.HostView(rng.integers(0, V, T).astype(np.int32), "i32", [T, 1]).expand([T, Dm])
    want = oracle.gather(table, ids, 0).numpy()
    got = H.download(B.gather(H.upload(ctx, table), H.upload(ctx, ids), 0))
    assert np.array_equal(got, want)
    g = H.HostView(rng.standard_normal(T * Dm).astype(np.float32), "f32", [T, Dm])
    zeros = H.HostView(np.zeros(V * Dm, np.float32), "f32", [V, Dm])
    want = oracle.scatter(zeros, ids, g, 0, "add").numpy()
    got = H.download(B.scatter(H.upload(ctx, zeros), H.upload(ctx, ids), H.upload(ctx, g), 0, "add"))
    H.assert_close("f32", got, want, rel=1e-5, abs_=1e-5, what="embedding backward")
    gi = H.HostView(rng.integers(-1000, 1000, T * Dm).astype(np.int32), "i32", [T, Dm])
    zi = H.HostView(np.zeros(V * Dm, np.int32), "i32", [V, Dm])
    want = oracle.scatter(zi, ids, gi, 0, "add").numpy()
    got = H.download(B.scatter(H.upload(ctx, zi), H.upload(ctx, ids), H.upload(ctx, gi), 0, "add"))
    assert np.array_equal(got, want)
    want = oracle.scatter(zi, ids, gi, 0, "set").numpy()  # many duplicates: last write must win
    got = H.download(B.scatter(H.upload(ctx, zi), H.upload(ctx, ids), H.upload(ctx, gi), 0, "set"))
    assert np.array_equal(got, want)


def test_threefry(ctx, oracle):
    def tf(k0, k1, c0, c1):
        key = H.HostView(np.array([k0, k1], dtype=np.uint32).view(np.int32), "i32", [1, 2])
        ctr = H.HostView(np.array([c0, c1], dtype=np.uint32).view(np.int32), "i32", [1, 2])
        return H.download(B.threefry(H.upload(ctx, key), H.upload(ctx, ctr))).reshape(-1).view(np.uint32).tolist()
    # Random123 known-answer vectors (backend_contract.ml:1663-1666)
    assert tf(0, 0, 0, 0) == [0x6b200159, 0x99ba4efe]
    assert tf(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == [0x1cb996fc, 0xbb002be7]
    assert tf(0x13198a2e, 0x03707344, 0x243f6a88, 0x85a308d3) == [0xc4923a9c, 0x483df7a0]
    rng = np.random.default_rng(2)
    n = 1 << 16
    key = H.HostView(rng.integers(-2**31, 2**31, 2 * n, dtype=np.int64).astype(np.int32), "i32", [2, n]).permute([1, 0])
    ctr = H.HostView(rng.integers(-2**31, 2**31, 2 * n, dtype=np.int64).astype(np.int32), "i32", [n, 2])
    want = oracle.threefry(key, ctr).numpy()
    got = H.download(B.threefry(H.upload(ctx, key), H.upload(ctx, ctr)))
    assert np.array_equal(got, want)
    z = H.HostView(np.zeros(6, np.int32), "i32", [2, 3])
    with pytest.raises(InvalidArgument, match="threefry: threefry: last axis must have extent 2"):
        B.threefry(H.upload(ctx, z), H.upload(ctx, z))


@pytest.mark.parametrize("dt", ["f32", "f64", "i32", "u8", "bf16", "i64", "bool", "c32"])
@pytest.mark.parametrize("op", ["sum", "prod", "max", "min"])
def test_scan(ctx, oracle, op, dt):
    rng = np.random.default_rng(8)
    cases = [hv for _, hv in H.layouts(dt, include_degenerate=False)]
    data = H.to_storage(dt, rng.uniform(0.9, 1.1, 64 * 300)) if dt in H.FLOATS else _rand(dt, 64 * 300, rng)
    big = H.HostView(data, dt, [64, 300])
    cases += [big, big.permute([1, 0]), big.flip([1])]
    for hv in cases:
        for axis in range(len(hv.shape)):
            try:
                want = oracle.scan(op, hv, axis).numpy()
            except Exception as e:
                with pytest.raises(Failure):
                    B.associative_scan(H.upload(ctx, hv), axis, op)
                return
            got = H.download(B.associative_scan(H.upload(ctx, hv), axis, op))
            if dt in H.COMPLEX:
                H.assert_close(dt, got, want, rel=1e-5, abs_=1e-5, what=f"cum{op}/{dt}")
            else:
                H.assert_same(dt, got, want, ulp=0, what=f"cum{op}/{dt}/{hv.shape}/{axis}")


def _dev_packed(ctx, hv):
    """Upload a packed HostView: storage bytes verbatim, nibble shape/offset on the handle."""
    from raven_b200 import dtype as D
    base = B.from_host(ctx, hv.storage, D.of(hv.dtype))
    return B.Tensor(base.buffer, hv.shape, hv.strides, hv.offset, base.dtype, ctx)


@pytest.mark.parametrize("pk", ["i4", "u4"])
def test_packed_int4(ctx, oracle, pk):
    """int4 / uint4 casts for every compute dtype at odd nibble offsets, packed->packed, and the
    odd-length prefix assign that must keep the neighbour nibble (backend_contract.ml:1402-1603)."""
    from raven_b200 import dtype as D
    from tests.test_gpu_map import _cast_inputs
    rng = np.random.default_rng(3)
    for dt in list(H.FLOATS) + list(H.INTS) + list(H.COMPLEX) + ["bool"]:
        data = _cast_inputs(dt)
        src = H.HostView(data.copy(), dt, [data.size])
        # compute -> packed into a pre-filled buffer at nibble offset 1: neighbours must survive
        want = H.HostView(np.full((data.size + 3) // 2 + 1, 0xA5, np.uint8), pk, [data.size], None, 1)
        oracle.assign  # noqa: B018  (oracle module present)
        import ctypes
        if oracle.__name__.endswith("ref"):
            oracle.call("cast", want, src)
        else:
            oracle._chk("cast", oracle.lib().nxo_cast(ctypes.byref(oracle._d(want)), ctypes.byref(oracle._d(src))))
        got_hv = H.HostView(np.full((data.size + 3) // 2 + 1, 0xA5, np.uint8), pk, [data.size], None, 1)
        out_t = _dev_packed(ctx, got_hv)
        do, ds = out_t._desc(), H.upload(ctx, src)._desc()
        st = ctx._lib.nxc_cast(ctx.ptr, ctypes.byref(do), ctypes.byref(ds))
        assert not st, st
        assert np.array_equal(B.to_host(out_t), want.storage), f"cast {dt}->{pk}"
        # packed -> compute from nibble offset 3
        packed = H.HostView(rng.integers(0, 256, 9).astype(np.uint8), pk, [15], None, 3)
        w = oracle.cast(packed, dt).numpy()
        g = H.download(B.cast(_dev_packed(ctx, packed), D.of(dt)))
        assert np.array_equal(H.raw(g), H.raw(w)), f"cast {pk}->{dt}"
    # odd-length prefix assign keeps the neighbour's high nibble
    base = _dev_packed(ctx, H.HostView(np.full(4, 0xFF, np.uint8), pk, [8]))
    prefix = B.Tensor(base.buffer, (5,), (1,), 0, base.dtype, ctx)
    B.assign(prefix, _dev_packed(ctx, H.HostView(np.array([0x21, 0x43, 0x05], np.uint8), pk, [5])))
    assert B.to_host(base).tolist() == [0x21, 0x43, 0xF5, 0xFF]
    strided = B.permute(_dev_packed(ctx, H.HostView(np.zeros(8, np.uint8), pk, [4, 4])), [1, 0])
    with pytest.raises(Failure, match="cast: packed dtype not supported for this operation"):
        B.cast(strided, DT.float32)
    with pytest.raises(Failure, match="add: packed dtype not supported for this operation"):
        B.add(base, base)


@pytest.mark.parametrize("dt", ["f32", "f64", "bf16", "f16", "i32", "u8", "i64", "u32", "bool", "c32"])
def test_sort_argsort(ctx, oracle, dt):
    """sort / argsort vs the oracle: NaN last both ways, stable argsort (bit-exact indices),
    complex lexicographic; short axes (several slices per shared-memory chunk), a 5000-long
    axis (global bitonic stages), strided and flipped inputs."""
    rng = np.random.default_rng(5)
    cases = [hv for _, hv in H.layouts(dt, include_degenerate=False)]
    n = 3 * 5000
    if dt in H.FLOATS:
        v = np.round(rng.uniform(-50, 50, n))
        v[rng.integers(0, n, 40)] = np.nan
        data = H.to_storage(dt, v)
    elif dt in H.COMPLEX:
        data = (np.round(rng.uniform(-3, 3, n)) + 1j * np.round(rng.uniform(-3, 3, n))).astype(H.np_storage(dt))
        data[[7, 4000]] = complex(np.nan, 1)
    elif dt == "bool":
        data = rng.integers(0, 2, n).astype(np.uint8)
    else:
        data = (_rand(dt, n, rng) % 97).astype(H.np_storage(dt))
    big = H.HostView(data, dt, [3, 5000])
    cases += [big, big.permute([1, 0]), big.flip([1]), H.HostView(data[:4096].copy(), dt, [2, 2048])]
    for hv in cases:
        for axis in range(len(hv.shape)):
            for desc in (False, True):
                want = oracle.argsort(hv, axis, desc).numpy()
                got = H.download(B.argsort(H.upload(ctx, hv), axis, desc))
                assert np.array_equal(got, want), f"argsort/{dt}/{hv.shape}/{axis}/{desc}"
                want = oracle.sort(hv, axis, desc).numpy()
                got = H.download(B.sort(H.upload(ctx, hv), axis, desc))
                if dt in H.COMPLEX:
                    assert np.array_equal(got, want, equal_nan=True) or np.array_equal(np.isnan(got), np.isnan(want))
                else:
                    assert np.array_equal(H.storage_to_float(dt, got), H.storage_to_float(dt, want), equal_nan=True), \
                        f"sort/{dt}/{hv.shape}/{axis}/{desc}"


WINDOW_CASES = [
    # (leading, spatial, kernel, stride, dilation, padding)
    ([2, 3], [7], [3], [1], [1], [(0, 0)]),
    ([2], [8], [3], [2], [2], [(1, 2)]),
    ([1, 2], [6, 5], [3, 2], [1, 1], [1, 1], [(0, 0), (0, 0)]),
    ([2, 2], [7, 6], [3, 3], [2, 1], [1, 2], [(1, 1), (2, 0)]),
    ([], [5, 4, 3], [2, 2, 2], [1, 2, 1], [1, 1, 1], [(0, 1), (1, 0), (0, 0)]),
    ([3], [4], [5], [1], [1], [(0, 0)]),
    ([4, 8], [32, 32], [3, 3], [1, 1], [1, 1], [(1, 1), (1, 1)]),   # a LeNet/ResNet-style 3x3 same conv
    ([2, 4], [28, 28], [5, 5], [2, 2], [1, 1], [(2, 2), (2, 2)]),
]


@pytest.mark.parametrize("dt", ["f32", "f64", "bf16", "f16", "f8e4m3", "i8", "i32", "u8", "i64", "u64", "bool", "c32", "c64"])
def test_unfold_fold(ctx, oracle, dt):
    """im2col / col2im vs the oracle, bit for bit (fold sums taps in the reference's order, so
    floats are exact too); contiguous and permuted inputs, padding, stride, dilation, K=1..3."""
    rng = np.random.default_rng(11)
    for lead, sp, k, s, d, p in WINDOW_CASES:
        shape = lead + sp
        n = int(np.prod(shape))
        x = H.HostView(_rand(dt, n, rng), dt, shape)
        want = oracle.unfold(x, k, s, d, p)
        got = B.unfold(H.upload(ctx, x), k, s, d, p)
        assert tuple(got.shape) == tuple(want.shape)
        assert np.array_equal(H.raw(H.download(got)), H.raw(want.numpy())), f"unfold/{dt}/{shape}/{k}"
        if want.shape[-1] > 0:
            w2 = oracle.fold(want, sp, k, s, d, p)
            g2 = B.fold(got, sp, k, s, d, p)
            assert np.array_equal(H.raw(H.download(g2)), H.raw(w2.numpy())), f"fold/{dt}/{shape}/{k}"
        if len(shape) >= 2:
            perm = list(range(len(shape)))[::-1]
            xp = H.HostView(_rand(dt, n, rng), dt, shape[::-1]).permute(perm)
            if list(xp.shape) == shape:
                want = oracle.unfold(xp, k, s, d, p)
                got = B.unfold(H.upload(ctx, xp), k, s, d, p)
                assert np.array_equal(H.raw(H.download(got)), H.raw(want.numpy())), f"unfold-strided/{dt}/{shape}"


def test_unfold_fold_errors(ctx):
    x = B.full(ctx, DT.float32, [2, 3, 8], 1.0)
    cols = B.unfold(x, [3], [1], [1], [(0, 0)])
    assert tuple(cols.shape) == (2, 3, 3, 6)
    with pytest.raises(InvalidArgument, match="fold: shape mismatch"):
        B.fold(cols, [9], [3], [1], [1], [(0, 0)])   # 9 - 3 + 1 = 7 windows, the columns hold 6
    p = B.buffer(ctx, DT.int4, [8])
    with pytest.raises(Failure, match="unfold: packed dtype"):
        B.unfold(p, [2], [1], [1], [(0, 0)])
