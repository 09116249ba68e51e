#!/usr/bin/env python
"""Timing experiment (not a bench): the C3 end-to-end leg (pinned host A and B through the C ABI) as a function of the
sample size, the H2D block size and the operator path."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402

torch.cuda.set_device(0)
d, n = 4096, 512
S = rb.DenseSkOp(rb.DenseDist(d, 4000000, rb.ScalarDist.Gaussian), rb.RNGState(1997), np.float64)
for mm in (50000, 100000, 200000):
    hA = torch.empty(mm * n, dtype=torch.float64, pin_memory=True)
    hA.normal_()
    hB = torch.zeros(d * n, dtype=torch.float64, pin_memory=True)
    for mat in (1, 0):
        for chunk in (16, 32, 64, 128):
            rb.set_option("dmma_materialise", mat)
            rb.set_option("h2d_chunk_mb", chunk)
            f = lambda: rb.sketch_general("C", "N", "N", d, n, mm, 1.0, S, 0, 0, hA.numpy(), mm, 0.0, hB.numpy(), d)
            f(); f()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                f()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 5
            print(f"m={mm} dmma_materialise={mat} h2d_chunk_mb={chunk}: {dt * 1e3:.2f} ms, {mm * n * 8 / 1e9 / dt:.2f} GB/s of A", flush=True)
