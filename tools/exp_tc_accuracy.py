#!/usr/bin/env python
"""Accuracy / time of the tcgen05 3xTF32 sketch at C1 against an fp64 product of the materialised operator (not a bench)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402
from exp_dense_layouts import timeit  # noqa: E402

torch.cuda.set_device(0)
d, m, n = 1024, 100000, 1024
S = rb.DenseSkOp(rb.DenseDist(d, m, rb.ScalarDist.Uniform), rb.RNGState(1997), np.float32)
A = torch.randn(m * n, dtype=torch.float32, device="cuda")
B = torch.zeros(d * n, dtype=torch.float32, device="cuda")
t = timeit(lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, B, d))
Sm = torch.empty(d * m, dtype=torch.float32, device="cuda")
rb.fill_dense(S.dist, Sm, rb.RNGState(1997))
want = (Sm.view(d, m).double() @ A.view(n, m).t().double())          # d x n
got = B.view(n, d).t().double()
err = float(torch.linalg.norm(got - want) / torch.linalg.norm(want))
print(f"C1: {t:.3f} ms, {2.0 * d * m * n / t / 1e9:.1f} TFLOP/s, relative Frobenius error vs fp64 {err:.2e}")
