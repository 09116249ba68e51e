#!/usr/bin/env python
"""Timing experiment (not a bench): sketch_general with a materialised operator (S.buff filled) on the tensor-core
kernels vs the fused (regenerating) path vs the SIMT generic kernel.  python tools/exp_materialised.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402


def timeit(f, reps=5):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(dt, d, m, n, fam, lay):
    tdt = torch.float32 if dt == np.float32 else torch.float64
    D = rb.DenseDist(d, m, fam, "L")
    S0 = rb.DenseSkOp(D, rb.RNGState(1997), dt)
    S1 = rb.DenseSkOp(D, rb.RNGState(1997), dt)
    rb.fill_dense(S1)
    A = torch.randn(m * n, dtype=tdt, device="cuda")
    B = torch.zeros(d * n, dtype=tdt, device="cuda")
    lda, ldb = (m, d) if lay == "C" else (n, n)
    res = {}
    for name, S, path in (("fused", S0, 0), ("materialised, tensor cores", S1, 0), ("materialised, generic SIMT", S1, 1)):
        rb.set_option("dense_path", path)
        ms = timeit(lambda: rb.sketch_general(lay, "N", "N", d, n, m, 1.0, S, 0, 0, A, lda, 0.0, B, ldb), 3 if path else 8)
        rb.set_option("dense_path", 0)
        res[name] = ms
        print(f"{np.dtype(dt).name} d={d} m={m} n={n} {fam} data {'K' if lay == 'C' else 'Q'}-contiguous  {name:28s} "
              f"{ms:9.3f} ms  {2.0 * d * m * n / ms / 1e9:8.1f} TFLOP/s", flush=True)
    return res


if __name__ == "__main__":
    torch.cuda.set_device(0)
    run(np.float32, 1024, 100000, 1024, "U", "C")
    run(np.float32, 1024, 100000, 1024, "G", "R")
    run(np.float64, 4096, 100000, 512, "G", "C")
    run(np.float64, 4096, 100000, 512, "G", "R")
