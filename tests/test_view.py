"""Host-side view algebra (raven_b200/view.py) against numpy's own strided views: same
shapes, strides, offsets and element order; same failures as the reference's View module
(packages/nx/lib/core/view.ml:82-322) for reshapes the strides cannot express."""
import numpy as np
import pytest

from raven_b200.view import View, c_contiguous_strides


def _materialize(v: View, storage):
    idx = np.full(v.shape, v.offset, dtype=np.int64)
    for ax, (s, st) in enumerate(zip(v.shape, v.strides)):
        sh = [1] * len(v.shape)
        sh[ax] = s
        idx = idx + (np.arange(s) * st).reshape(sh)
    return storage[idx.reshape(-1)].reshape(v.shape)


def test_movement_ops_match_numpy():
    base = np.arange(2 * 3 * 4 * 5)
    a = base.reshape(2, 3, 4, 5)
    v = View((2, 3, 4, 5))
    assert np.array_equal(_materialize(v.permute([3, 1, 0, 2]), base), a.transpose(3, 1, 0, 2))
    assert np.array_equal(_materialize(v.shrink([(0, 2), (1, 3), (0, 4), (2, 5)]), base), a[:, 1:3, :, 2:5])
    assert np.array_equal(_materialize(v.flip([True, False, True, False]), base), a[::-1, :, ::-1, :])
    assert np.array_equal(_materialize(v.reshape((6, 20)), base), a.reshape(6, 20))
    assert np.array_equal(_materialize(View((1, 5)).expand((7, 5)), base), np.broadcast_to(base[:5], (7, 5)))
    assert View(()).expand((2, 2)).strides == (0, 0)
    # reshape of strided views: split / merge where strides compose, size-1 insertions
    t = v.permute([1, 0, 2, 3])
    assert np.array_equal(_materialize(t.reshape((3, 2, 20)), base), a.transpose(1, 0, 2, 3).reshape(3, 2, 20))
    assert np.array_equal(_materialize(t.reshape((3, 1, 2, 2, 2, 5)), base),
                          a.transpose(1, 0, 2, 3).reshape(3, 1, 2, 2, 2, 5))
    s = v.shrink([(0, 2), (0, 3), (0, 4), (0, 3)])
    assert np.array_equal(_materialize(s.reshape((6, 4, 3)), base), a[..., :3].reshape(6, 4, 3))
    with pytest.raises(ValueError, match="call contiguous"):
        t.reshape((6, 20))
    with pytest.raises(ValueError, match="cannot reshape"):
        v.reshape((7, 7))
    with pytest.raises(ValueError, match="only singletons expand"):
        v.expand((2, 3, 4, 6))
    with pytest.raises(ValueError, match="duplicate axis"):
        v.permute([0, 0, 1, 2])
    with pytest.raises(ValueError, match="bounds must be within shape"):
        v.shrink([(0, 3), (0, 3), (0, 4), (0, 5)])


def test_contiguous_strides_follow_the_reference_zero_extent_rule():
    assert c_contiguous_strides((2, 3, 4)) == [12, 4, 1]
    assert c_contiguous_strides((0, 4)) == [0, 1]       # core/shape.ml:22-33
    assert c_contiguous_strides((3, 0, 2)) == [0, 0, 1]
    assert View((0, 4)).offset == 0 and View((3, 4)).reshape((0, 7)).shape == (0, 7)
