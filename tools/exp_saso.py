#!/usr/bin/env python
"""Timing experiment for the SASO apply (not a bench): python tools/exp_saso.py [m]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402


def run(path, d, m, n, k, reps=5, dt=np.float32):
    rb.set_option("saso_path", path)
    tdt = torch.float32 if dt == np.float32 else torch.float64
    torch.manual_seed(0)                     # the same A for every path
    S = rb.SparseSkOp(rb.SparseDist(d, m, k), rb.RNGState(1997), dtype=dt)
    A = torch.randn(m * n, dtype=tdt, device="cuda")
    B = torch.zeros(d * n, dtype=tdt, device="cuda")
    f = lambda: rb.sketch_general("R", "N", "N", d, n, m, 1.0, S, 0, 0, A, n, 0.0, B, n)
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{np.dtype(dt).name} path={path} d={d} m={m} n={n} k={k}: {ms:.3f} ms, {m * n * np.dtype(dt).itemsize / ms / 1e6:.1f} GB/s of A", flush=True)
    return B


if __name__ == "__main__":
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 8000000
    torch.cuda.set_device(0)
    B2 = run(2, 2048, m, 256, 8)
    B1 = run(1, 2048, m, 256, 8, reps=2)
    print("binned vs atomic kernel rel diff", float(torch.linalg.norm(B2 - B1) / torch.linalg.norm(B1)))
    D2 = run(2, 2048, m // 2, 256, 8, dt=np.float64)          # same bytes of A as the float problem
    D1 = run(1, 2048, m // 2, 256, 8, reps=2, dt=np.float64)
    print("double: binned vs atomic kernel rel diff", float(torch.linalg.norm(D2 - D1) / torch.linalg.norm(D1)))
    run(2, 2048, m, 256, 4)
    run(2, 1024, m, 256, 8)
    run(2, 2048, m // 2, 512, 8)
    rb.set_option("saso_path", 0)
