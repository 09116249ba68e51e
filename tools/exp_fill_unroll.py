#!/usr/bin/env python
"""Timing experiment (not a bench): Uniform float fill_dense of long vectors, 4 against 16 Philox blocks per thread and tile
(fill_unroll = 0 / 1); the outputs are compared bit for bit."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
import randblas_b200 as rb
def timeit(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
rows, cols = 2048, 1000000
D = rb.DenseDist(8192, cols, rb.ScalarDist.Uniform, rb.Axis.Long)
buf = torch.empty(rows * cols, dtype=torch.float32, device="cuda")
ref = None
for v in (0, 1, 0, 1):
    rb.set_option("fill_unroll", v)
    t = timeit(lambda: rb.fill_dense_unpacked("R", D, rows, cols, 0, 0, buf, rb.RNGState(1997)))
    gs = rows * cols / t / 1e6
    if ref is None: ref = buf.clone()
    print(f"fill_unroll={v}: {t:.3f} ms, {gs:.0f} Gsamples/s, {gs*4/1e3:.2f} TB/s, equal to first: {bool(torch.equal(ref, buf))}", flush=True)
