"""Multi-rank correctness worker: one process per GPU (launched by tests/test_gpu_multirank.py
through torch.distributed.run, or by hand with torchrun on a multi-GPU box).

The reference's own data-parallel tests compare the multi-device result with the single-device
one on the same inputs (reference: packages/kaun/test/test_pmap_dp.ml:18); so does this: every
rank builds the SAME full array from a shared seed, keeps its leading-axis slab, runs the sharded
path (raven_b200.sharded: local kernel + the peer-memory / NCCL exchange) and compares with the
single-GPU answer of the same backend over the full array -- exact for integers, bools and
indices, within the dtype's tolerance for float sums (one more combine level, rounded once more
through the storage type) -- and with the reference oracle where that is cheap. Results must also
be BIT-IDENTICAL across ranks.

Exit code 0 = every check passed on this rank; rank 0 prints a one-line summary."""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as td  # noqa: E402

import raven_b200.backend as B  # noqa: E402
from raven_b200 import dtype as D  # noqa: E402
from raven_b200 import sharded  # noqa: E402
from tests import harness as H  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = B.create_context(device=local)


def _exchange(idbytes):
    t = torch.tensor(list(idbytes), dtype=torch.uint8, device="cuda")
    td.broadcast(t, 0)
    return bytes(t.cpu().tolist())


comm = sharded.NcclComm(ctx, rank, world, _exchange)
P2P = bool(ctx._lib.nxc_dist_p2p_enabled(ctx.ptr))
checks = 0


def same_on_all_ranks(arr: np.ndarray, what: str):
    digest = hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()
    got = [None] * world
    td.all_gather_object(got, digest)
    assert len(set(got)) == 1, f"{what}: results differ across ranks: {got}"


def full_and_slab(dtype, rows_per_rank, cols, rng, lo=-4.0, hi=4.0):
    n = world * rows_per_rank * cols
    if dtype in H.FLOATS:
        vals = rng.uniform(lo, hi, n)
    elif dtype in H.COMPLEX:
        vals = rng.uniform(lo, hi, n) + 1j * rng.uniform(lo, hi, n)
    elif dtype == "bool":
        vals = rng.integers(0, 2, n)
    elif dtype in H.UINTS:
        vals = rng.integers(0, 200, n)
    else:
        vals = rng.integers(-100, 100, n)
    st = H.to_storage(dtype, vals)
    full = H.HostView(st, dtype, [world * rows_per_rank, cols])
    slab = full.shrink([(rank * rows_per_rank, (rank + 1) * rows_per_rank), (0, cols)])
    return full, slab


def tol_for(dtype, op):
    if dtype in H.INTS or dtype == "bool":
        return None
    if op in ("max", "min"):
        return 0.0
    return {"f64": 1e-12, "c64": 1e-12, "f32": 1e-5, "c32": 1e-5, "f16": 4e-3, "bf16": 3e-2, "f8e4m3": 0.3, "f8e5m2": 0.6}[dtype]


def check_reduce():
    global checks
    rng = np.random.default_rng(1234)
    for dtype in ("f32", "f64", "i32", "i16", "u16", "u8", "i64", "bf16", "f16", "f8e4m3", "c32", "bool"):
        for op in ("sum", "prod", "max", "min"):
            if dtype == "bool" and op in ("sum", "prod"):
                continue
            if dtype in H.COMPLEX and op in ("max", "min"):
                continue
            for axes, rows, cols in (([0], 6, 40), ([0, 1], 6, 40), ([1], 6, 40), ([0], 3, 70000), ([0], 2, 200000)):
                # 70000 / 200000 columns: partials above the 256 KiB mailbox slot -> NCCL where it has the
                # reduction (f32 / i32 sum), all-gather + the backend's own fold where it has not (int16,
                # NaN-sticky float max)
                if cols == 70000 and (dtype not in ("f32", "i32") or op not in ("sum", "max")):
                    continue
                if cols == 200000 and (dtype, op) not in (("i16", "sum"), ("i16", "max"), ("bf16", "max"), ("f32", "min")):
                    continue
                span = (0.6, 1.4) if op == "prod" else (-4.0, 4.0)
                full, slab = full_and_slab(dtype, rows, cols, rng, *span)
                got = H.download(sharded.sharded_reduce(H.upload(ctx, slab), op, axes, comm))
                want = H.download(B.reduce(H.upload(ctx, full), op, axes))
                what = f"sharded_reduce {op} {dtype} axes={axes} [{rows}x{cols} per rank]"
                tol = tol_for(dtype, op)
                if tol is None or tol == 0.0:
                    H.assert_same(dtype, got, want, ulp=0, what=what)
                else:
                    scale = float(np.max(np.abs(H.storage_to_float(dtype, want)))) if dtype not in H.COMPLEX else float(np.max(np.abs(want)))
                    H.assert_close(dtype, got, want, rel=tol, abs_=tol * max(scale, 1.0), what=what)
                if 0 in axes:
                    same_on_all_ranks(got, what)
                checks += 1
    # NaN is sticky through the exchange (NCCL's max / min would drop it): rank 1 holds the NaN
    full, slab = full_and_slab("f32", 4, 33, rng)
    st = full.storage.copy()
    st[(1 % world) * 4 * 33 + 5] = np.nan
    full = H.HostView(st, "f32", [world * 4, 33])
    slab = full.shrink([(rank * 4, (rank + 1) * 4), (0, 33)])
    for op in ("max", "min"):
        got = H.download(sharded.sharded_reduce(H.upload(ctx, slab), op, [0], comm))
        want = H.download(B.reduce(H.upload(ctx, full), op, [0]))
        assert np.isnan(got[5]) and np.isnan(want[5]), f"NaN lost in sharded {op}: {got[5]} / {want[5]}"
        H.assert_same("f32", got, want, ulp=0, what=f"sharded {op} with a NaN on rank 1")
        checks += 1


def check_argreduce():
    global checks
    rng = np.random.default_rng(4321)
    oracle = H.get_oracle()
    for dtype in ("f32", "f64", "i32", "u8", "bf16", "i64"):
        for rows, cols in ((5, 1), (7, 33), (3, 3000), (2, 20000)):   # 20000 outputs: past the fused path's limit
            full, slab = full_and_slab(dtype, rows, cols, rng)
            st = full.storage.copy()
            flat_cols = cols
            # ties across ranks (first index must win), the extreme on the LAST rank, a NaN on rank 1
            if cols >= 33:
                st[0 * flat_cols + 3] = st[(world * rows - 1) * flat_cols + 3] = H.to_storage(dtype, np.array([100]))[0]
                st[(world * rows - 1) * flat_cols + 4] = H.to_storage(dtype, np.array([101]))[0]
                if dtype in H.FLOATS:
                    nanbits = H.to_storage(dtype, np.array([np.nan]))[0]
                    st[((1 % world) * rows + 1) * flat_cols + 6] = nanbits
                    st[((world - 1) * rows) * flat_cols + 6] = nanbits
            full = H.HostView(st, dtype, [world * rows, cols])
            slab = full.shrink([(rank * rows, (rank + 1) * rows), (0, cols)])
            for is_max in (True, False):
                name = "argmax" if is_max else "argmin"
                got = H.download(sharded.sharded_argreduce(H.upload(ctx, slab), is_max, 0, rank * rows, comm))
                want = H.download((B.argmax if is_max else B.argmin)(H.upload(ctx, full), 0))
                H.assert_same("i32", got, want, what=f"sharded {name} {dtype} [{rows}x{cols} per rank]")
                if cols <= 3000:
                    ref = oracle.argreduce(name, full, 0).numpy()
                    H.assert_same("i32", got, ref, what=f"sharded {name} {dtype} vs oracle")
                same_on_all_ranks(got, name)
                checks += 1
    # along the other axis: no exchange but an all-gather of disjoint outputs
    full, slab = full_and_slab("f32", 4, 50, rng)
    got = H.download(sharded.sharded_argreduce(H.upload(ctx, slab), True, 1, 0, comm))
    want = H.download(B.argmax(H.upload(ctx, full), 1))
    H.assert_same("i32", got, want, what="sharded argmax along the unsharded axis")
    # a 1-D slab, as bench.py's step does it
    n = 1 << 16
    v = np.random.default_rng(99).uniform(-1, 1, world * n).astype(np.float32)
    v[(world - 1) * n + 17] = 7.0
    full = H.HostView(v, "f32", [world * n])
    slab = full.shrink([(rank * n, (rank + 1) * n)])
    got = H.download(sharded.sharded_argreduce(H.upload(ctx, slab), True, 0, rank * n, comm))
    assert int(got) == (world - 1) * n + 17, f"1-D sharded argmax: {int(got)}"
    checks += 2


def check_batch_matmul_and_dp():
    global checks
    rng = np.random.default_rng(77)
    bpr, m, k, n = 3, 64, 96, 128
    a = rng.standard_normal((world * bpr, m, k)).astype(np.float32)
    b = rng.standard_normal((world * bpr, k, n)).astype(np.float32)
    for dtype in ("f32", "bf16"):
        fa = H.HostView(H.to_storage(dtype, a).reshape(-1), dtype, [world * bpr, m, k])
        fb = H.HostView(H.to_storage(dtype, b).reshape(-1), dtype, [world * bpr, k, n])
        sl = [(rank * bpr, (rank + 1) * bpr)]
        got = H.download(sharded.sharded_batch_matmul(H.upload(ctx, fa.shrink(sl + [(0, m), (0, k)])),
                                                      H.upload(ctx, fb.shrink(sl + [(0, k), (0, n)])), comm, gather=True))
        want = H.download(B.matmul(H.upload(ctx, fa), H.upload(ctx, fb)))
        H.assert_same(dtype, got, want, ulp=0, what=f"sharded batch matmul {dtype} (gathered) vs single GPU")
        same_on_all_ranks(got, "batch matmul")
        checks += 1
    # gradient averaging: every rank contributes rank-dependent leaves; mean must equal the closed form
    leaves = [B.full(ctx, D.float32, [1000], float(rank + 1)), B.full(ctx, D.float32, [300, 1000], 0.5 * (rank + 1)),
              B.full(ctx, D.bfloat16, [4096], float(rank + 1))]
    mean = (world + 1) / 2.0
    for out, scale, dtype in zip(sharded.allreduce_mean_(leaves, comm), (1.0, 0.5, 1.0), ("f32", "f32", "bf16")):
        got = H.storage_to_float(dtype, H.download(out))
        assert np.allclose(got, mean * scale, rtol=1e-2 if dtype == "bf16" else 1e-6), f"allreduce_mean_ {dtype}: {got.flat[:3]}"
        checks += 1
    red = sharded.GradBucketReducer(comm)
    for t in (B.full(ctx, D.float32, [1 << 20], float(rank + 1)), B.full(ctx, D.float32, [77], 2.0 * (rank + 1))):
        red.push(t)
    for out, scale in zip(red.finish(), (1.0, 2.0)):
        assert np.allclose(H.download(out), mean * scale, rtol=1e-6), "GradBucketReducer"
        checks += 1


def check_capture():
    """A captured sharded step (local kernels + the exchange kernels, whose epoch lives on the
    device) replayed several times over refreshed inputs gives the eager answers."""
    global checks
    n = 1 << 18
    rng = np.random.default_rng(5 + rank)
    x = B.from_host(ctx, rng.uniform(-1, 1, n).astype(np.float32))
    X = B.reshape(x, [256, n // 256])

    def step():
        return (sharded.sharded_reduce(x, "sum", [0], comm), sharded.sharded_reduce(X, "max", [0], comm),
                sharded.sharded_argreduce(x, True, 0, rank * n, comm))

    eager = [H.download(t) for t in step()]
    with ctx.capture() as g:
        outs = step()
    assert g.kernels >= 3
    for it in range(3):
        g.launch()
        got = [H.download(t) for t in outs]
        for a, b in zip(got, eager):
            assert np.array_equal(a, b), f"replay {it} differs from the eager step"
        # eager exchanges interleave with replays (same sequence on every rank)
        again = [H.download(t) for t in step()]
        for a, b in zip(again, eager):
            assert np.array_equal(a, b)
    # refreshed input: plant a new maximum on the last rank, in place
    v = rng.uniform(-1, 1, n).astype(np.float32)
    if rank == world - 1:
        v[123] = 9.0
    B.assign(x, B.from_host(ctx, v))
    g.launch()
    assert int(H.download(outs[2])) == (world - 1) * n + 123, "replay does not see the refreshed input"
    g.close()
    checks += 1


def check_dp_training():
    """Kaun-style data parallelism end to end (reference: packages/kaun/test/test_pmap_dp.ml:18 compares
    the multi-device trajectory with the single-device one): a tiny GPT-2 trained for a few SGD steps
    with the batch sharded over the ranks and gradients averaged through FlatBucketReducer must follow
    the single-GPU run on the whole batch -- losses and final parameters, to f32 reduction order."""
    global checks
    from tools import gpt2_step as G
    cfg, per_rank, seq = G.GPT2_TINY, 2, 12
    grid = np.random.default_rng(321).integers(0, cfg["vocab"], (per_rank * world, seq + 1))
    host = G.init_params_host(cfg, seed=11)
    dp = G.Trainer(B, ctx, cfg, per_rank, seq, opt="sgd", lr=5e-2, comm=comm, bucket_mb=0, host_params={k: v.copy() for k, v in host.items()})
    dp.bucket_bytes = 16 << 10   # several buckets even at this size
    mine = grid[rank * per_rank:(rank + 1) * per_rank]
    dp.set_batch(mine[:, :-1], mine[:, 1:])
    one = G.Trainer(B, ctx, cfg, per_rank * world, seq, opt="sgd", lr=5e-2, comm=None, host_params={k: v.copy() for k, v in host.items()})
    one.set_batch(grid[:, :-1], grid[:, 1:])
    for step in range(4):
        dp.pre_step()
        one.pre_step()
        l_dp = sharded.allreduce_mean_([dp.step_body()], comm)[0]
        l_one = one.step_body()
        a, b = float(H.download(l_dp)), float(H.download(l_one))
        assert abs(a - b) <= 1e-5 * abs(b), f"step {step}: data-parallel loss {a} vs single-GPU {b}"
    for k in one.params:
        w, g = H.download(one.params[k]).astype(np.float64), H.download(dp.params[k]).astype(np.float64)
        assert np.abs(w - g).max() <= 1e-5 * (np.abs(w).max() + 1e-3), f"parameter {k} diverged"
        same_on_all_ranks(H.download(dp.params[k]), k)
    checks += 1


try:
    check_reduce()
    check_dp_training()
    check_argreduce()
    check_batch_matmul_and_dp()
    check_capture()
    ctx.sync()
    total = torch.tensor([checks], device="cuda")
    td.all_reduce(total)
    if rank == 0:
        print(f"multirank ok: world={world} p2p={P2P} checks/rank={checks} total={int(total.item())}")
finally:
    try:
        comm.close()
    except Exception:
        pass
    td.destroy_process_group()
