"""Pins the oracle before anything trusts it.

oracle/nxo.c (the C restatement) is compared, op by op, with the reference's own
C backend compiled unmodified from /root/reference (oracle/_ref/libnxref.so) over
the reference contract suite's pools and layout matrix
(packages/nx/test/backend_contract.ml:125-289, 423-468, 555-593): every unary,
binary and comparison op x every dtype x every layout (including which
(op, dtype) pairs are rejected, and with which exception class and message),
where, the 17 x 17 cast matrix with the contract's edge cases, copy/assign,
reductions over axes [0] / [last] / all, argmax/argmin with keepdims, matmul
over transposed / batched / broadcast operands.

Bit-exact everywhere except: float sum/prod (reassociation; 4 ulp of the compute
type) and complex transcendentals (libm vs libm: identical here, both are glibc).
Skipped when oracle/_ref has not been built (it needs /root/reference).
"""
import numpy as np
import pytest

from oracle import nxo, ref
from tests import harness as H

pytestmark = pytest.mark.skipif(not (ref.available() and nxo.available()),
                                reason="oracle/_ref/libnxref.so or oracle/libnxo.so not built")

ALL = list(H.FLOATS) + list(H.INTS) + list(H.COMPLEX) + ["bool"]


def _both(fn):
    out = []
    for m in (ref, nxo):
        try:
            out.append(fn(m))
        except (ref.RefError, nxo.RefError) as e:
            out.append(("err", e.kind, e.msg))
    return out


def _same(dt, r, o, what, ulp=0):
    if isinstance(r, tuple) or isinstance(o, tuple):
        assert r == o, f"{what}: reference {r} vs restatement {o}"
        return False
    x, y = o.numpy(), r.numpy()
    if np.array_equal(H.raw(x), H.raw(y)):
        return True
    H.assert_same(dt, x, y, ulp=ulp, what=what)  # NaN payload / sign-of-zero differences only
    return True


@pytest.mark.parametrize("dt", ALL)
def test_unary_matches_reference(dt):
    for op in ref.UNARY:
        for name, hv in H.layouts(dt):
            r, o = _both(lambda m: m.unary(op, hv))
            if not _same(dt, r, o, f"{op}/{dt}/{name}"):
                break


@pytest.mark.parametrize("dt", ALL)
def test_binary_and_compare_match_reference(dt):
    la, lb = dict(H.layouts(dt)), dict(H.layouts(dt, rot=5))
    for op in ref.BINARY + ["shl", "shr"]:
        for na, nb in H.BINARY_LAYOUT_PAIRS:
            r, o = _both(lambda m: m.binary(op, la[na], lb[nb]))
            if not _same(dt, r, o, f"{op}/{dt}/{na},{nb}"):
                break
    for op in ref.CMP:
        for na, nb in H.BINARY_LAYOUT_PAIRS:
            r, o = _both(lambda m: m.compare(op, la[na], lb[nb]))
            if not _same("bool", r, o, f"{op}/{dt}/{na},{nb}"):
                break


@pytest.mark.parametrize("dt", ALL)
def test_where_copy_assign_match_reference(dt):
    la, lb, lc = dict(H.layouts(dt)), dict(H.layouts(dt, rot=7)), dict(H.layouts("bool", rot=2))
    for nc, na, nb in [("contig", "contig", "contig"), ("transpose", "transpose", "transpose"),
                       ("broadcast", "flip", "slice"), ("scalar", "scalar", "scalar"), ("empty", "empty", "empty")]:
        r, o = _both(lambda m: m.where(lc[nc], la[na], lb[nb]))
        _same(dt, r, o, f"where/{dt}/{nc}")
    for name, hv in H.layouts(dt):
        r, o = _both(lambda m: m.copy(hv))
        _same(dt, r, o, f"copy/{dt}/{name}")
    res = []
    for m in (ref, nxo):
        base = H.HostView(H.pool(dt, 18).copy(), dt, [3, 6])
        m.assign(base.shrink([(0, 3), (1, 5)]).flip([0]), la["transpose"])
        res.append(base.storage.copy())
    assert np.array_equal(H.raw(res[0]), H.raw(res[1]))


@pytest.mark.parametrize("src", ALL)
def test_cast_matrix_matches_reference(src):
    from tests.test_gpu_map import _cast_inputs
    data = _cast_inputs(src)
    n = (data.size // 2) * 2
    for dst in ALL:
        for hv in (H.HostView(data.copy(), src, [data.size]),
                   H.HostView(data[:n].copy(), src, [n // 2, 2]).permute([1, 0])):
            r, o = _both(lambda m: m.cast(hv, dst))
            _same(dst, r, o, f"cast {src}->{dst}")


@pytest.mark.parametrize("dt", ALL)
def test_reduce_and_argreduce_match_reference(dt):
    for op in ("sum", "prod", "max", "min"):
        for name, hv in H.layouts(dt):
            nd = len(hv.shape)
            for axes in ([[]] if nd == 0 else [[0], [nd - 1], list(range(nd))]):
                r, o = _both(lambda m: m.reduce(op, hv, axes))
                exact = dt in H.INTS or dt == "bool" or op in ("max", "min")
                _same(dt, r, o, f"{op}/{dt}/{name}/{axes}", ulp=0 if exact else 4)
    if dt in H.COMPLEX:
        return
    for op in ("argmax", "argmin"):
        for name, hv in H.layouts(dt, include_degenerate=False):
            for axis in (0, len(hv.shape) - 1):
                for keep in (False, True):
                    r, o = _both(lambda m: m.argreduce(op, hv, axis, keep))
                    _same("i32", r, o, f"{op}/{dt}/{name}/{axis}")
    for op in ("sum", "prod", "max", "min"):
        for name, hv in H.layouts(dt, include_degenerate=False):
            r, o = _both(lambda m: m.scan(op, hv, len(hv.shape) - 1))
            exact = dt in H.INTS or dt == "bool" or op in ("max", "min")
            _same(dt, r, o, f"cum{op}/{dt}/{name}", ulp=0 if exact else 4)


@pytest.mark.parametrize("dt", [d for d in ALL if d != "bool"])
def test_matmul_matches_reference(dt):
    from tests.test_gpu_matmul import _fill
    for (m, k, n) in [(5, 7, 3), (1, 1, 1), (16, 9, 24)]:
        A, B = _fill(dt, m, k, n)
        a = H.HostView(A.reshape(-1).copy(), dt, [m, k])
        b = H.HostView(B.reshape(-1).copy(), dt, [k, n])
        at = H.HostView(np.ascontiguousarray(A.T).reshape(-1), dt, [k, m]).permute([1, 0])
        bt = H.HostView(np.ascontiguousarray(B.T).reshape(-1), dt, [n, k]).permute([1, 0])
        for x, y in ((a, b), (at, b), (a, bt), (at, bt)):
            r, o = _both(lambda mod: mod.matmul(x, y))
            if dt in H.INTS:
                _same(dt, r, o, f"matmul/{dt}")
            else:
                H.assert_close(dt, o.numpy(), r.numpy(), rel=1e-5 if dt not in ("f64", "c64") else 1e-12,
                               abs_=1e-5 if dt not in ("f64", "c64") else 1e-12, what=f"matmul/{dt}")


def test_large_random_arrays_match_reference():
    rng = np.random.default_rng(1)
    n = 1 << 16
    x = H.HostView(rng.uniform(-20, 20, n).astype(np.float32), "f32", [n])
    y = H.HostView(rng.uniform(-20, 20, n).astype(np.float32), "f32", [n])
    for op in ref.UNARY:
        r, o = _both(lambda m: m.unary(op, x))
        _same("f32", r, o, op)
    for op in ("add", "mul", "fdiv", "pow", "atan2", "mod", "idiv", "max"):
        r, o = _both(lambda m: m.binary(op, x, y))
        _same("f32", r, o, op)
    m2 = H.HostView(x.storage, "f32", [256, 256])
    for axes in ([0], [1], [0, 1]):
        r, o = _both(lambda m: m.reduce("sum", m2, axes))
        H.assert_close("f32", o.numpy(), r.numpy(), rel=2e-6, abs_=1e-4, what=f"sum {axes}")
        r, o = _both(lambda m: m.reduce("max", m2.permute([1, 0]), axes))
        _same("f32", r, o, f"max {axes}")
        if len(axes) == 1:
            r, o = _both(lambda m: m.argreduce("argmin", m2.flip([0]), axes[0]))
            _same("i32", r, o, f"argmin {axes}")


def test_move_and_random_match_reference():
    """pad / cat / gather / scatter / threefry, including the
    Random123 known-answer vectors the contract pins (backend_contract.ml:1663-1666)."""
    rng = np.random.default_rng(4)
    for dt in ("f32", "i32", "u8", "bf16", "c32", "bool", "i64"):
        base = H.HostView(H.pool(dt, 18).copy(), dt, [3, 6])
        views = {"contig": base, "T": H.HostView(H.pool(dt, 18).copy(), dt, [6, 3]).permute([1, 0]),
                 "slice": H.HostView(np.tile(H.pool(dt, 18), 2).copy(), dt, [3, 12]).shrink([(0, 3), (2, 8)])}
        fill = H.HostView(H.pool(dt, 18)[4:5].copy(), dt, [])
        for name, v in views.items():
            r, o = _both(lambda m: m.pad(v, [(1, 2), (0, 3)], fill))
            _same(dt, r, o, f"pad/{dt}/{name}")
            r, o = _both(lambda m: m.cat([v, views["contig"], v], 0))
            _same(dt, r, o, f"cat0/{dt}/{name}")
            r, o = _both(lambda m: m.cat([v, views["T"]], 1))
            _same(dt, r, o, f"cat1/{dt}/{name}")
            for axis, hi in ((0, 3), (1, 6)):
                idx = H.HostView(rng.integers(-hi, hi, 4 * 5).astype(np.int32), "i32", [4, 5])
                shape = [4, 6] if axis == 0 else [3, 5]
                idx = H.HostView(rng.integers(-hi, hi, shape[0] * shape[1]).astype(np.int32), "i32", shape)
                r, o = _both(lambda m: m.gather(v, idx, axis))
                _same(dt, r, o, f"gather/{dt}/{name}/{axis}")
                upd = H.HostView(np.resize(H.pool(dt, 18), shape[0] * shape[1]).copy(), dt, shape)
                for mode in ("set", "add"):
                    r, o = _both(lambda m: m.scatter(v, idx, upd, axis, mode))
                    _same(dt, r, o, f"scatter-{mode}/{dt}/{name}/{axis}")
        bad = H.HostView(np.array([0, 7, 1], dtype=np.int32), "i32", [1, 3]).expand([3, 3])
        r, o = _both(lambda m: m.gather(base, H.HostView(np.array([0, 9, 1] * 6, dtype=np.int32), "i32", [3, 6]), 0))
        assert isinstance(r, tuple) and r == o and r[1] == "Failure"
    # threefry: the three Random123 KATs + random keys/counters over strided operands
    def tf(m, k0, k1, c0, c1):
        key = H.HostView(np.array([k0, k1], dtype=np.uint32).view(np.int32), "i32", [1, 2])
        ctr = H.HostView(np.array([c0, c1], dtype=np.uint32).view(np.int32), "i32", [1, 2])
        return m.threefry(key, ctr).numpy().reshape(-1).view(np.uint32).tolist()
    for m in (ref, nxo):
        assert tf(m, 0, 0, 0, 0) == [0x6b200159, 0x99ba4efe]
        assert tf(m, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == [0x1cb996fc, 0xbb002be7]
        assert tf(m, 0x13198a2e, 0x03707344, 0x243f6a88, 0x85a308d3) == [0xc4923a9c, 0x483df7a0]
    key = H.HostView(rng.integers(-2**31, 2**31, 64, dtype=np.int64).astype(np.int32), "i32", [2, 32]).permute([1, 0])
    ctr = H.HostView(rng.integers(-2**31, 2**31, 64, dtype=np.int64).astype(np.int32), "i32", [32, 2])
    r, o = _both(lambda m: m.threefry(key, ctr))
    _same("i32", r, o, "threefry")
    r, o = _both(lambda m: m.threefry(H.HostView(np.zeros(6, np.int32), "i32", [2, 3]), H.HostView(np.zeros(6, np.int32), "i32", [2, 3])))
    assert isinstance(r, tuple) and r == o and r[1] == "Invalid_argument", (r, o)


def test_packed_int4_matches_reference():
    """int4 / uint4: cast both ways for every compute dtype, nibble offsets, packed->packed,
    and the odd-length prefix assign that must keep the neighbour nibble
    (backend_contract.ml:1402-1513, 1517-1603)."""
    rng = np.random.default_rng(3)
    for pk in ("i4", "u4"):
        for dt in ALL:
            from tests.test_gpu_map import _cast_inputs
            data = _cast_inputs(dt)
            src = H.HostView(data.copy(), dt, [data.size])
            outs = []
            for m in (ref, nxo):
                out = H.HostView(np.full((data.size + 3) // 2 + 1, 0xA5, np.uint8), pk, [data.size], None, 1)
                m.call("cast", out, src) if m is ref else nxo._chk("cast", nxo.lib().nxo_cast(
                    __import__("ctypes").byref(nxo._d(out)), __import__("ctypes").byref(nxo._d(src))))
                outs.append(out.storage.copy())
            assert np.array_equal(outs[0], outs[1]), f"cast {dt}->{pk}"
            packed = H.HostView(rng.integers(0, 256, 9).astype(np.uint8), pk, [15], None, 3)
            r, o = _both(lambda m: m.cast(packed, dt))
            _same(dt, r, o, f"cast {pk}->{dt}")
        a = H.HostView(rng.integers(0, 256, 8).astype(np.uint8), pk, [13], None, 2)
        r, o = _both(lambda m: m.cast(a, "u4" if pk == "i4" else "i4"))
        assert np.array_equal(r.storage, o.storage)
        res = []
        for m in (ref, nxo):
            base = H.HostView(np.full(4, 0xFF, np.uint8), pk, [8])
            prefix = H.HostView(base.storage, pk, [5])
            m.assign(prefix, H.HostView(np.array([0x21, 0x43, 0x05], np.uint8), pk, [5]))
            res.append(base.storage.copy())
        assert np.array_equal(res[0], res[1]) and res[0].tolist() == [0x21, 0x43, 0xF5, 0xFF]
        strided = H.HostView(np.zeros(8, np.uint8), pk, [4, 4]).permute([1, 0])
        r, o = _both(lambda m: m.cast(strided, "f32"))
        assert isinstance(r, tuple) and r == o and "packed" in r[2]


@pytest.mark.parametrize("dt", ALL)
def test_sort_argsort_match_reference(dt):
    """sort / argsort: NaN last in both directions, stable argsort, complex lexicographic
    (nx_c_sort.c). Value sort is compared numerically (equal keys -- +0/-0, NaN payloads --
    are interchangeable there, as the reference documents); argsort exactly."""
    rng = np.random.default_rng(5)
    cases = [hv for _, hv in H.layouts(dt, include_degenerate=False)]
    if dt in H.FLOATS:
        v = np.round(rng.uniform(-5, 5, 200))
        v[[3, 50, 120]] = np.nan
        cases.append(H.HostView(H.to_storage(dt, v), dt, [10, 20]))
    elif dt in H.COMPLEX:
        v = (np.round(rng.uniform(-3, 3, 200)) + 1j * np.round(rng.uniform(-3, 3, 200))).astype(H.np_storage(dt))
        v[7] = complex(np.nan, 1)
        cases.append(H.HostView(v, dt, [10, 20]))
    else:
        from tests.test_gpu_map import _rand
        cases.append(H.HostView(_rand(dt, 200, rng) if dt == "bool" else (_rand(dt, 200, rng) % 7).astype(H.np_storage(dt)), dt, [10, 20]))
    for hv in cases:
        for axis in range(len(hv.shape)):
            for desc in (False, True):
                r, o = _both(lambda m: m.argsort(hv, axis, desc))
                _same("i32", r, o, f"argsort/{dt}/{hv.shape}/{axis}/{desc}")
                r, o = _both(lambda m: m.sort(hv, axis, desc))
                x, y = o.numpy(), r.numpy()
                if dt in H.COMPLEX:
                    assert np.array_equal(x, y, equal_nan=True) or np.array_equal(np.isnan(x), np.isnan(y))
                else:
                    fx, fy = H.storage_to_float(dt, x), H.storage_to_float(dt, y)
                    assert np.array_equal(fx, fy, equal_nan=True), f"sort/{dt}/{hv.shape}/{axis}/{desc}"


WINDOW_CASES = [
    # (leading, spatial, kernel, stride, dilation, padding)
    ([2, 3], [7], [3], [1], [1], [(0, 0)]),
    ([2], [8], [3], [2], [2], [(1, 2)]),
    ([1, 2], [6, 5], [3, 2], [1, 1], [1, 1], [(0, 0), (0, 0)]),
    ([2, 2], [7, 6], [3, 3], [2, 1], [1, 2], [(1, 1), (2, 0)]),
    ([], [5, 4, 3], [2, 2, 2], [1, 2, 1], [1, 1, 1], [(0, 1), (1, 0), (0, 0)]),
    ([3], [4], [5], [1], [1], [(0, 0)]),   # kernel larger than the extent: win clamps to 1
]


@pytest.mark.parametrize("dt", ALL)
def test_unfold_fold_match_reference(dt):
    """im2col / col2im (nx_c_move.c:588-870): padded taps read zero; fold sums overlapping taps in
    the compute type in ascending kernel-offset order -- compared bit for bit, floats included."""
    from tests.test_gpu_map import _rand
    rng = np.random.default_rng(11)
    for lead, sp, k, s, d, p in WINDOW_CASES:
        shape = lead + sp
        n = int(np.prod(shape))
        x = H.HostView(_rand(dt, n, rng), dt, shape)
        r, o = _both(lambda m: m.unfold(x, k, s, d, p))
        _same(dt, r, o, f"unfold/{dt}/{shape}/{k}")
        cols = o
        if cols.shape[-1] > 0:  # L == 0 with a clamped window count is undefined in the reference
            r, o = _both(lambda m: m.fold(cols, sp, k, s, d, p))
            _same(dt, r, o, f"fold/{dt}/{shape}/{k}")
        if len(shape) >= 2:  # a permuted (strided) input
            xp = H.HostView(_rand(dt, n, rng), dt, shape[::-1]).permute(list(range(len(shape)))[::-1])
            if list(xp.shape) == shape:
                r, o = _both(lambda m: m.unfold(xp, k, s, d, p))
                _same(dt, r, o, f"unfold-strided/{dt}/{shape}/{k}")
    x = H.HostView(_rand(dt, 0, rng), dt, [0, 5])
    r, o = _both(lambda m: m.unfold(x, [2], [1], [1], [(0, 0)]))
    _same(dt, r, o, "unfold-empty")


def test_unfold_fold_packed_rejected():
    x = H.HostView(np.zeros(4, np.uint8), "i4", [8])
    r, o = _both(lambda m: m.unfold(x, [2], [1], [1], [(0, 0)]))
    assert isinstance(r, tuple) and r == o
