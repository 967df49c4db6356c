#!/usr/bin/env python
"""One eigvals call on a 512 x 512 f32 matrix: the process tools/ncu_one.sh profiles for the eig kernel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raven_b200.backend as B  # noqa: E402

ctx = B.create_context(device=0)
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
a = rng.standard_normal((1, n, n)).astype(np.float32)
ta = B.reshape(B.from_host(ctx, a.reshape(-1)), [1, n, n])
B.eigvals(ta)
ctx.sync()
