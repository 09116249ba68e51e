#!/usr/bin/env python
"""Timing experiment for sketch_sparse (not a bench): python tools/exp_sksp.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402


def run(d, m, n, per_row, reps=3):
    lens = torch.poisson(torch.full((m,), per_row, device="cuda")).to(torch.int64)
    rowptr = torch.zeros(m + 1, dtype=torch.int64, device="cuda")
    torch.cumsum(lens, 0, out=rowptr[1:])
    nnz = int(rowptr[-1].item())
    col = torch.randint(0, n, (nnz,), device="cuda", dtype=torch.int64)
    vals = torch.randn(nnz, device="cuda", dtype=torch.float32)
    A = rb.CSRMatrix(m, n, nnz, vals, rowptr, col)
    S = rb.DenseSkOp(rb.DenseDist(d, 10000000), rb.RNGState(1997), np.float32)
    B = torch.zeros(d * n, dtype=torch.float32, device="cuda")
    f = lambda: rb.sketch_sparse("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, 0.0, B, d)
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"d={d} m={m} n={n} per_row={per_row} nnz={nnz} B={d*n*4/1e6:.0f}MB: {ms:.2f} ms, "
          f"{ms*1e6/ (nnz*d/4):.3f} ns per red.v4, {nnz*d/4/ms/1e6:.1f} G red.v4/s", flush=True)


if __name__ == "__main__":
    torch.cuda.set_device(0)
    cfgs = [(512, 2000000, 125000, 12.5), (512, 2000000, 31250, 12.5), (512, 2000000, 15625, 12.5),
            (128, 2000000, 125000, 12.5), (128, 2000000, 62500, 12.5), (128, 2000000, 31250, 12.5),
            (256, 2000000, 62500, 12.5), (512, 2000000, 125000, 50.0), (512, 500000, 125000, 100.0)]
    for c in cfgs:
        run(*c)
