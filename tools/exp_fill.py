#!/usr/bin/env python
"""Timing experiment (not a bench): fill_dense for both families and both scalar types, natural and transposed layout."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402
from exp_dense_layouts import timeit  # noqa: E402

torch.cuda.set_device(0)
rows, cols = 2048, 1000000
for fam, fname in ((rb.ScalarDist.Gaussian, "Gaussian"), (rb.ScalarDist.Uniform, "Uniform")):
    for dt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
        D = rb.DenseDist(8192, cols, fam, rb.Axis.Long)
        buf = torch.empty(rows * cols, dtype=tdt, device="cuda")
        for lay in ("R", "C"):
            t = timeit(lambda: rb.fill_dense_unpacked(lay, D, rows, cols, 0, 0, buf, rb.RNGState(1997)), reps=3)
            gs = rows * cols / t / 1e6
            print(f"{fname} {np.dtype(dt).name} layout {lay}: {t:.3f} ms, {gs:.0f} Gsamples/s, "
                  f"{gs * np.dtype(dt).itemsize / 1e3:.2f} TB/s written", flush=True)
        del buf
