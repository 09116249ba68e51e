#!/usr/bin/env python
"""Timing experiment (not a bench): float tensor-core sketch of Q-contiguous data, tiles fed to the tensor core as an MN-major
operand (tc_ymn = 0) against the transposing path (tc_ymn = 1) and against K-contiguous data; results compared with each other
and with an fp64 product of the materialised operator."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402
from exp_pair import timeit  # noqa: E402

torch.cuda.set_device(0)
d, m, n = 1024, 100000, 1024
flops = 2.0 * d * m * n
A = torch.randn(m * n, dtype=torch.float32, device="cuda")
for fam in (rb.ScalarDist.Uniform, rb.ScalarDist.Gaussian):
    S = rb.DenseSkOp(rb.DenseDist(d, m, fam), rb.RNGState(1997), np.float32)
    Sd = torch.empty(d * m, dtype=torch.float64, device="cuda")
    rb.fill_dense_unpacked("R", S.dist, d, m, 0, 0, Sd, S.seed_state)
    want = (Sd.view(d, m) @ A.view(m, n).double())                       # RowMajor A (m x n): Q-contiguous
    B = torch.zeros(d * n, dtype=torch.float32, device="cuda")
    t = timeit(lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, B, d))
    print(f"{fam} K-contiguous data: {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s", flush=True)
    out = {}
    for mode in (0, 1):
        rb.set_option("tc_ymn", mode)
        B = torch.zeros(d * n, dtype=torch.float32, device="cuda")
        before = rb.counter("tensor_core_launches")
        t = timeit(lambda: rb.sketch_general("R", "N", "N", d, n, m, 1.0, S, 0, 0, A, n, 0.0, B, n))
        out[mode] = B.clone()
        err = float((B.view(d, n).double() - want).norm() / want.norm())
        print(f"{fam} Q-contiguous data, tc_ymn={mode}: {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s, rel err vs fp64 {err:.2e}, "
              f"tensor-core launches {rb.counter('tensor_core_launches') - before}", flush=True)
    print("   bit-identical:", bool(torch.equal(out[0], out[1])), flush=True)
rb.set_option("tc_ymn", 0)
# right sketch of ColMajor data (the range-finder call A * S), ragged shape
mm, dd, nn = 100000, 1000, 1000
St = rb.DenseSkOp(rb.DenseDist(mm, dd, rb.ScalarDist.Uniform), rb.RNGState(7), np.float32)
A2 = torch.randn(nn * mm, dtype=torch.float32, device="cuda")           # ColMajor n x m, lda = n
for mode in (0, 1):
    rb.set_option("tc_ymn", mode)
    B = torch.zeros(nn * dd, dtype=torch.float32, device="cuda")
    t = timeit(lambda: rb.sketch_general("C", "N", "N", nn, dd, mm, 1.0, A2, nn, St, 0, 0, 0.0, B, nn))
    print(f"right sketch A(n x m) S(m x d), ColMajor, n = d = 1000, tc_ymn={mode}: {t:.3f} ms", flush=True)
rb.set_option("tc_ymn", 0)
