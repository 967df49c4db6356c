#!/usr/bin/env python
"""One bf16 8192^3 product per layout (NN, NT, TN) after a warm-up each, for ncu:
   ncu --set full --clock-control none -k regex:nxc_mm_tc -o /tmp/mm_lay python tools/mm_layouts_once.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raven_b200.backend as B  # noqa: E402
from raven_b200 import dtype as D  # noqa: E402

ctx = B.create_context(device=0)
M = int(os.environ.get("MM_M", 8192))
h = np.random.default_rng(0).uniform(-1, 1, 1 << 22).astype(np.float32)
t = B.from_host(ctx, h)
t = B.reshape(B.contiguous(B.expand(B.reshape(t, [1, 1 << 22]), [M * M >> 22, 1 << 22])), [M, M])
x = B.cast(t, D.bfloat16)
y = B.cast(t, D.bfloat16)
for lay, (p, q) in {"NN": (x, y), "NT": (x, B.permute(y, [1, 0])), "TN": (B.permute(x, [1, 0]), y)}.items():
    B.matmul(p, q)
    ctx.sync()
