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
