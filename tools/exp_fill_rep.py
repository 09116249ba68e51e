#!/usr/bin/env python
"""Timing experiment (not a bench): Gaussian fill_dense at the C2 shape (a 1024 x 1e6 window of the 8192 x 1e6 operator), one
pass of 8 Philox blocks per thread and tile (fill_rep = 0) against tiles of 4 / 2 passes (1 / 2); outputs compared bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402


def timeit(f, reps=5):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rows, cols = 1024, 1000000
for dt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
    D = rb.DenseDist(8192, cols, rb.ScalarDist.Gaussian, rb.Axis.Long)
    buf = torch.empty(rows * cols, dtype=tdt, device="cuda")
    ref = None
    for v in (0, 1, 2, 0, 1, 2):
        rb.set_option("fill_rep", v)
        t = timeit(lambda: rb.fill_dense_unpacked("R", D, rows, cols, 0, 0, buf, rb.RNGState(1997)))
        gs = rows * cols / t / 1e6
        if ref is None:
            ref = buf.clone()
        print(f"{np.dtype(dt).name} fill_rep={v}: {t:.3f} ms, {gs:.0f} Gsamples/s, {gs * buf.element_size() / 1e3:.2f} TB/s, "
              f"equal to the first: {bool(torch.equal(ref, buf))}", flush=True)
    del buf, ref
