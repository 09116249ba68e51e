#!/usr/bin/env python
"""Timing experiment (not a bench): DMMA dense sketch, Gaussian vs Uniform operator, C3 shard shape and a smaller one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402
from exp_dense_layouts import timeit  # noqa: E402

torch.cuda.set_device(0)
for (d, m, n) in ((4096, 500000, 512), (1024, 50000, 512)):
    A = torch.randn(m * n, dtype=torch.float64, device="cuda")
    B = torch.zeros(d * n, dtype=torch.float64, device="cuda")
    for fam in (rb.ScalarDist.Gaussian, rb.ScalarDist.Uniform):
        S = rb.DenseSkOp(rb.DenseDist(d, m, fam), rb.RNGState(1997), np.float64)
        for uni in (0, 1):
            rb.set_option("dmma_uniform_warps", uni)
            t = timeit(lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, B, d), reps=2)
            print(f"d={d} m={m} n={n} {fam} {'uniform warps' if uni else 'warp-specialised'}: {t:.3f} ms, "
                  f"{2.0 * d * m * n / t / 1e9:.1f} TFLOP/s", flush=True)
        rb.set_option("dmma_uniform_warps", 0)
    del A, B
