"""GPU tests of the runtime around the kernels: the copy engines (pinned uploads on their own
stream, asynchronous read-back with deferred frees) and stream ordering. The reference has no
counterpart (its buffers are host memory, backend_c/nx_backend.ml:50-69); what must hold is that
from_host / to_host keep their copy semantics whatever stream the bytes travel on."""
import numpy as np
import pytest

import raven_b200.backend as B
from raven_b200 import Failure
from raven_b200 import dtype as DT

pytestmark = pytest.mark.gpu


def test_pinned_upload_engine(ctx):
    n = 1 << 21   # 8 MiB: above the 1 MiB threshold of the upload engine
    pa, pb = ctx.pinned_empty(n, np.float32), ctx.pinned_empty(n, np.float32)
    rng = np.random.default_rng(0)
    pa[:] = rng.standard_normal(n)
    pb[:] = rng.standard_normal(n)
    for _ in range(3):   # uploads interleaved with compute that reads them
        a, b = B.from_host(ctx, pa), B.from_host(ctx, pb)
        c = B.add(a, b)
        assert np.array_equal(B.to_host(c), pa + pb)
    small = ctx.pinned_empty(16, np.int32)   # below the threshold: context-stream path
    small[:] = np.arange(16)
    assert np.array_equal(B.to_host(B.from_host(ctx, small)), np.arange(16, dtype=np.int32))


def test_async_readback_defers_free(ctx):
    n = 1 << 22   # 16 MiB
    pa = ctx.pinned_empty(n, np.float32)
    pa[:] = np.arange(n, dtype=np.float32) % 1000
    a = B.from_host(ctx, pa)
    outs = [ctx.pinned_empty(n, np.float32) for _ in range(4)]
    for k, out in enumerate(outs):
        t = B.add(a, B.expand(B.full(ctx, DT.float32, [], float(k + 1)), [n]))
        B.to_host_async(t, out)
        del t   # freed while the copy may still be in flight: the engine must keep it alive
        # churn the allocator with same-sized buffers full of a poison value
        for _ in range(3):
            junk = B.expand(B.full(ctx, DT.float32, [], -7.0), [n])
            junk = B.contiguous(junk)
            del junk
    ctx.sync()
    for k, out in enumerate(outs):
        assert np.array_equal(out, pa + np.float32(k + 1)), f"read-back {k} was overwritten"


def test_async_readback_needs_pinned(ctx):
    t = B.full(ctx, DT.float32, [1 << 10], 1.0)
    with pytest.raises(Failure, match="pinned"):
        B.to_host_async(t, np.empty(1 << 10, np.float32))


def test_two_readbacks_of_one_buffer_then_drop(ctx):
    """Two asynchronous read-backs of the SAME tensor, then the tensor is dropped: the buffer must
    outlive the LAST copy, not the first (the engine releases it with the last pending entry)."""
    n = 1 << 23   # 32 MiB: the second copy is still queued when the first finishes
    a = B.contiguous(B.expand(B.full(ctx, DT.float32, [], 3.5), [n]))
    o1, o2 = ctx.pinned_empty(n, np.float32), ctx.pinned_empty(n, np.float32)
    B.to_host_async(a, o1)
    B.to_host_async(a, o2)
    del a
    for _ in range(6):   # churn: same-sized buffers of another value, allocated while the copies run
        junk = B.contiguous(B.expand(B.full(ctx, DT.float32, [], -1.0), [n]))
        del junk
    ctx.sync()
    assert (o1 == 3.5).all() and (o2 == 3.5).all()


def test_unchecked_gather_reports_at_the_next_sync(ctx):
    """nxc_gather_trusted does not drain the stream; an out-of-range index is never dereferenced and
    its flag is sticky in the status page until nxc_sync / nxc_d2h reports it -- once."""
    data = B.full(ctx, DT.float32, [8, 4], 2.0)
    idx = B.reshape(B.from_host(ctx, np.array([0, 1, 99, 3, 4, 5, 6, 7], dtype=np.int32)), [2, 4])
    out = B.gather(data, idx, 0, trusted=True)
    with pytest.raises(Failure, match="index out of bounds"):
        ctx.sync()
    ctx.sync()
    assert B.to_numpy(out).shape == (2, 4)
