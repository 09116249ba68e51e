#!/usr/bin/env python
"""Timing experiment (not a bench): the double-precision tensor-core sketch at the C3 shard's shape: Gaussian operator fused
and through a generated panel (dmma_materialise), Uniform operator, materialised operator, both data layouts."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402
from exp_pair import timeit  # noqa: E402

torch.cuda.set_device(0)
d, n, m = 4096, 512, 500000
A = torch.randn(m * n, dtype=torch.float64, device="cuda")
flops = 2.0 * d * m * n
ref = {}
for fam, mats in ((rb.ScalarDist.Gaussian, (0, 1)), (rb.ScalarDist.Uniform, (0,))):
    for layout, lda, ldb in (("C", m, d), ("R", n, n)):
        S = rb.DenseSkOp(rb.DenseDist(d, 4000000, fam), rb.RNGState(1997), np.float64)
        out = {}
        for mat in mats:
            rb.set_option("dmma_materialise", mat)
            B = torch.zeros(d * n, dtype=torch.float64, device="cuda")
            t = timeit(lambda: rb.sketch_general(layout, "N", "N", d, n, m, 1.0, S, 0, 0, A, lda, 0.0, B, ldb), reps=3)
            out[mat] = B.clone()
            print(f"{fam} layout {layout} dmma_materialise={mat}: {t:.3f} ms, {flops / t / 1e9:.2f} TFLOP/s", flush=True)
        if len(out) == 2:
            print(f"   panel vs fused: rel diff {float((out[0] - out[1]).norm() / out[0].norm()):.2e}", flush=True)
rb.set_option("dmma_materialise", int(os.environ.get("DMMA_MAT_DEFAULT", "0")))
mm = 100000
S = rb.DenseSkOp(rb.DenseDist(d, mm), rb.RNGState(5), np.float64)
rb.fill_dense(S)
B = torch.zeros(d * n, dtype=torch.float64, device="cuda")
t = timeit(lambda: rb.sketch_general("C", "N", "N", d, n, mm, 1.0, S, 0, 0, A, mm, 0.0, B, d), reps=3)
print(f"materialised operator m={mm}: {t:.3f} ms, {2.0 * d * mm * n / t / 1e9:.2f} TFLOP/s", flush=True)
X = torch.randn(d, 32768, dtype=torch.float64, device="cuda")
Y = torch.randn(32768, n, dtype=torch.float64, device="cuda")
t = timeit(lambda: torch.mm(X, Y), reps=5)
print(f"cuBLAS DGEMM {d} x 32768 x {n}: {t:.3f} ms, {2.0 * d * 32768 * n / t / 1e9:.2f} TFLOP/s", flush=True)
# panel size sweep (Gaussian, layout C)
rb.set_option("dmma_materialise", 1)
S = rb.DenseSkOp(rb.DenseDist(d, 4000000, rb.ScalarDist.Gaussian), rb.RNGState(1997), np.float64)
for mb in (256, 512, 1024, 2048, 4096):
    rb.set_option("dmma_panel_mb", mb)
    B = torch.zeros(d * n, dtype=torch.float64, device="cuda")
    t = timeit(lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, B, d), reps=3)
    print(f"Gaussian, panel of {mb} MB: {t:.3f} ms, {flops / t / 1e9:.2f} TFLOP/s", flush=True)
rb.set_option("dmma_panel_mb", 2048)
