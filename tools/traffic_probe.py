#!/usr/bin/env python
"""The process bench.py puts under one `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`
pass to MEASURE the DRAM traffic of its dominant kernel in the same run (roofline.traffic):
three flat f32 adds over 2^log2n elements through the C ABI, nothing else named KBin."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raven_b200.backend as B  # noqa: E402
from raven_b200 import dtype as D  # noqa: E402

n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
ctx = B.create_context(device=int(os.environ.get("LOCAL_RANK", "0")))
a = B.full(ctx, D.float32, [n], 1.5)
b = B.full(ctx, D.float32, [n], 2.25)
for _ in range(3):
    c = B.add(a, b)
ctx.sync()
