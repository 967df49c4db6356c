import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import raven_b200.backend as B
from raven_b200 import dtype as D
ctx = B.create_context()
ctx.set_matmul_mode("tf32")
rng = np.random.default_rng(0)
def run(tag, A, Bm, ta=False, tb=False):
    m, k = A.shape; n = Bm.shape[1]
    if ta:
        a = B.permute(B.reshape(B.from_host(ctx, np.ascontiguousarray(A.T).reshape(-1)), [k, m]), [1, 0])
    else:
        a = B.reshape(B.from_host(ctx, A.reshape(-1)), [m, k])
    if tb:
        b = B.permute(B.reshape(B.from_host(ctx, np.ascontiguousarray(Bm.T).reshape(-1)), [n, k]), [1, 0])
    else:
        b = B.reshape(B.from_host(ctx, Bm.reshape(-1)), [k, n])
    c = B.to_numpy(B.matmul(a, b)).reshape(m, n)
    ref = A.astype(np.float64) @ Bm.astype(np.float64)
    err = np.abs(c - ref)
    print(tag, "maxerr/scale %.3g" % (err.max() / np.abs(ref).max()), "bad frac %.3f" % np.mean(err > 5e-3 * np.abs(ref).max()), flush=True)
for mode in ("0", "1"):
    os.environ["NX_CUDA_MM_PAIR"] = mode
    for (m, k, n) in [(128, 32, 256), (256, 64, 256), (512, 384, 640), (1024, 1024, 1024)]:
        A = rng.standard_normal((m, k)).astype(np.float32)
        Bm = rng.standard_normal((k, n)).astype(np.float32)
        for ta in (False, True):
            for tb in (False, True):
                run(f"pair={mode} {m}x{k}x{n} A{'mn' if ta else 'k'} B{'k' if tb else 'mn'}", A, Bm, ta, tb)
