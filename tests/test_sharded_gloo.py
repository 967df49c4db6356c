"""The N>1 host logic on CPU: world_size-2 `gloo` run of raven_b200.sharded (leading-axis
slabs -> local reduce -> one exchange step) with an oracle-backed backend double standing in
for the CUDA engine and a gloo-backed comm standing in for NCCL. What is checked is the
sharding arithmetic: slab bounds, global-index offsets, which collective each reduction
uses, NaN / tie resolution across ranks, and that the answer equals the single-device one
(bit-exact for ints, argmax/argmin and float max/min; reassociation tolerance for float sums).
"""
import os
import socket

import numpy as np
import pytest

from oracle.hostview import HostView
from raven_b200 import dtype as D
from raven_b200 import sharded


class OT:
    """Tensor double: a HostView with the attributes sharded.py reads."""

    def __init__(self, hv):
        self.hv, self.dtype, self.shape, self.context = hv, D.of(hv.dtype), tuple(hv.shape), None


class OracleBackend:
    """Duck-typed stand-in for raven_b200.backend, computing with the oracle (tests only)."""

    def __init__(self):
        from tests import harness as H
        self.o = H.get_oracle()

    def reduce(self, x, op, axes):
        return OT(self.o.reduce(op, x.hv, axes))

    def argmax(self, x, axis, keepdims=False):
        return OT(self.o.argreduce("argmax", x.hv, axis, keepdims))

    def argmin(self, x, axis, keepdims=False):
        return OT(self.o.argreduce("argmin", x.hv, axis, keepdims))

    def add(self, a, b):
        return OT(self.o.binary("add", a.hv, b.hv))

    def mul(self, a, b):
        return OT(self.o.binary("mul", a.hv, b.hv))

    def full(self, ctx, dt, shape, value):
        n = int(np.prod(shape)) if len(shape) else 1
        return OT(HostView(np.full(n, value, dtype=D.of(dt).np), D.of(dt).name, shape))

    def expand(self, t, shape):
        return OT(t.hv.expand(shape))

    def reshape(self, t, shape):
        return OT(self.o.copy(t.hv).reshape_contig(shape))

    def contiguous(self, t):
        return OT(self.o.copy(t.hv))

    def gather(self, data, idx, axis):
        return OT(self.o.gather(data.hv, idx.hv, axis))

    def matmul(self, a, b):
        return OT(self.o.matmul(a.hv, b.hv))


class GlooComm:
    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def allreduce(self, t, op):
        import torch
        import torch.distributed as td
        arr = t.hv.numpy().copy()
        ten = torch.from_numpy(arr.reshape(-1))
        td.all_reduce(ten, op={"sum": td.ReduceOp.SUM, "prod": td.ReduceOp.PRODUCT, "max": td.ReduceOp.MAX,
                               "min": td.ReduceOp.MIN}[op])
        return OT(HostView(ten.numpy().copy(), t.hv.dtype, t.shape))

    def allgather(self, t):
        import torch
        import torch.distributed as td
        arr = np.ascontiguousarray(t.hv.numpy())
        ten = torch.from_numpy(arr.reshape(-1).copy())
        outs = [torch.empty_like(ten) for _ in range(self.world)]
        td.all_gather(outs, ten)
        flat = np.concatenate([o.numpy() for o in outs])
        return OT(HostView(flat, t.hv.dtype, (self.world,) + tuple(t.shape)))


def _worker(rank, world, port, q):
    try:
        import torch.distributed as td
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        td.init_process_group("gloo", rank=rank, world_size=world)
        be, comm = OracleBackend(), GlooComm(rank, world)
        o = be.o
        rng = np.random.default_rng(0)  # same stream on every rank: the full array is known everywhere
        R, C = 38, 24
        full_f = rng.uniform(-1, 1, R * C).astype(np.float32)
        full_f[5 * C + 3] = np.nan          # a NaN in rank 0's slab
        full_f[30 * C + 3] = np.nan         # and one in rank 1's, same column: rank 0's must win argmax
        full_f[7 * C + 9] = 9.0
        full_f[33 * C + 9] = 9.0            # a tie across ranks: the lower global index wins
        full_i = rng.integers(-2**31, 2**31, R * C, dtype=np.int64).astype(np.int32)
        lo, hi = sharded.slab_bounds(R, rank, world)
        assert sharded.slab_bounds(37, 0, 2) == (0, 19) and sharded.slab_bounds(37, 1, 2) == (19, 37)
        assert (lo, hi) == ((0, 19) if rank == 0 else (19, 38))
        for name, full in (("f32", full_f), ("i32", full_i)):
            whole = HostView(full.copy(), name, [R, C])
            mine = OT(HostView(full[lo * C:hi * C].copy(), name, [hi - lo, C]))
            for op in ("sum", "max", "min", "prod"):
                if op == "prod" and name == "f32":
                    continue
                for axes in ([0], [0, 1], [1]):
                    got = sharded.sharded_reduce(mine, op, axes, comm, backend=be).hv.numpy()
                    want = o.reduce(op, whole, axes).numpy()
                    if name == "i32" or op in ("max", "min"):
                        assert np.array_equal(got, want, equal_nan=True), (name, op, axes)
                    else:
                        assert np.allclose(got, want, rtol=1e-5, atol=1e-5, equal_nan=True), (name, op, axes)
            for is_max in (True, False):
                got = sharded.sharded_argreduce(mine, is_max, 0, lo, comm, backend=be).hv.numpy()
                want = o.argreduce("argmax" if is_max else "argmin", whole, 0).numpy()
                assert np.array_equal(got, want), (name, is_max, got, want)
                got = sharded.sharded_argreduce(mine, is_max, 1, lo, comm, backend=be).hv.numpy()
                want = o.argreduce("argmax" if is_max else "argmin", whole, 1).numpy()
                assert np.array_equal(got, want), (name, is_max, "axis1")
        # batch-leading matmul: independent slabs, allgather of C on request
        A = rng.standard_normal((4, 5, 6)).astype(np.float32)
        Bm = rng.standard_normal((4, 6, 3)).astype(np.float32)
        blo, bhi = sharded.slab_bounds(4, rank, world)
        c = sharded.sharded_batch_matmul(OT(HostView(A[blo:bhi].reshape(-1).copy(), "f32", [bhi - blo, 5, 6])),
                                         OT(HostView(Bm[blo:bhi].reshape(-1).copy(), "f32", [bhi - blo, 6, 3])),
                                         comm, gather=True, backend=be).hv.numpy()
        assert np.allclose(c, A @ Bm, rtol=1e-5, atol=1e-5)
        # data-parallel gradient averaging
        g = OT(HostView(np.full(8, float(rank + 1), np.float32), "f32", [8]))
        avg = sharded.allreduce_mean_([g], comm, backend=be)[0].hv.numpy()
        assert np.allclose(avg, 1.5)
        # the bucketed reducer (async on NCCL, synchronous on this double) gives the same leaves
        red = sharded.GradBucketReducer(comm, backend=be)
        for k in range(3):
            red.push(OT(HostView(np.full(4 + k, float((rank + 1) * (k + 1)), np.float32), "f32", [4 + k])))
        outs = red.finish()
        assert [tuple(t.hv.shape) for t in outs] == [(4,), (5,), (6,)]
        for k, t in enumerate(outs):
            assert np.allclose(t.hv.numpy(), 1.5 * (k + 1))
        td.barrier()
        td.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


def test_sharded_paths_world2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(timeout=30)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
