#!/usr/bin/env python
"""A handful of bf16 8192^3 products (for `ncu -k regex:nxc_mm_tc_kernel -s 2 -c 1`) and, with
argv[1] == fold, inner-axis sums of 2-byte and 1-byte [16384, 16384] arrays."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raven_b200.backend as B  # noqa: E402

ctx = B.create_context(device=0)
n = 8192
a = B.cast(B.reshape(B.from_host(ctx, np.random.default_rng(0).standard_normal(n * n).astype(np.float32)), [n, n]), "bf16")
b = B.cast(B.reshape(B.from_host(ctx, np.random.default_rng(1).standard_normal(n * n).astype(np.float32)), [n, n]), "bf16")
if len(sys.argv) > 1 and sys.argv[1] == "fold":
    big = B.reshape(B.contiguous(B.expand(B.reshape(a, [1, n * n]), [4, n * n])), [16384, 16384])
    for _ in range(3):
        B.reduce(big, "sum", [1])
        B.reduce(B.cast(big, "i8"), "sum", [1])
else:
    for _ in range(4):
        c = B.matmul(a, b)
ctx.sync()
