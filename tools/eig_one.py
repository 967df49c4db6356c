import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import raven_b200.backend as B
ctx = B.create_context(device=0)
rng = np.random.default_rng(0)
n = 512
a = rng.standard_normal((1, n, n)).astype(np.float32)
ta = B.reshape(B.from_host(ctx, a.reshape(-1)), [1, n, n])
B.eigvals(ta)
ctx.sync()
