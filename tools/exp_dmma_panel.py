#!/usr/bin/env python
"""Timing experiment (not a bench): double Gaussian sketch at the C3 shape, fused against panel-materialise + XMAT."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402
from exp_pair import timeit  # noqa: E402

torch.cuda.set_device(0)
d, n = 4096, 512
for m in (100000, 500000):
    A = torch.randn(m * n, dtype=torch.float64, device="cuda")
    S = rb.DenseSkOp(rb.DenseDist(d, 4000000, rb.ScalarDist.Gaussian), rb.RNGState(1997), np.float64)
    flops = 2.0 * d * m * n
    out = {}
    for mat in (0, 1):
        rb.set_option("dmma_materialise", mat)
        B = torch.zeros(d * n, dtype=torch.float64, device="cuda")
        t = timeit(lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, B, d), reps=3)
        out[mat] = B.clone()
        print(f"m={m} dmma_materialise={mat}: {t:.3f} ms, {flops / t / 1e9:.2f} TFLOP/s", flush=True)
    rel = float((out[0] - out[1]).norm() / out[0].norm())
    print(f"   rel diff {rel:.2e}", flush=True)
    del A
rb.set_option("dmma_materialise", 0)
