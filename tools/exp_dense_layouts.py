#!/usr/bin/env python
"""Timing experiment (not a bench): float dense sketch with K-contiguous vs Q-contiguous data, tensor-core vs generic."""
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


def main():
    torch.cuda.set_device(0)
    for dt, tdt in ((np.float32, torch.float32), (np.float64, torch.float64)):
        d, m, n = (1024, 100000, 1024) if dt == np.float32 else (1024, 50000, 512)
        S = rb.DenseSkOp(rb.DenseDist(d, m, rb.ScalarDist.Uniform), rb.RNGState(1997), dt)
        A = torch.randn(m * n, dtype=tdt, device="cuda")
        B = torch.zeros(d * n, dtype=tdt, device="cuda")
        flops = 2.0 * d * m * n
        t = timeit(lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, B, d))
        print(f"{np.dtype(dt).name} left ColMajor (K-contiguous): {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s", flush=True)
        t = timeit(lambda: rb.sketch_general("R", "N", "N", d, n, m, 1.0, S, 0, 0, A, n, 0.0, B, n))
        print(f"{np.dtype(dt).name} left RowMajor (Q-contiguous): {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s", flush=True)
        St = rb.DenseSkOp(rb.DenseDist(m, d, rb.ScalarDist.Uniform), rb.RNGState(1997), dt)
        t = timeit(lambda: rb.sketch_general("C", "N", "N", n, d, m, 1.0, A, n, St, 0, 0, 0.0, B, n))
        print(f"{np.dtype(dt).name} right ColMajor A(n x m) * S(m x d) (Q-contiguous): {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s", flush=True)
        # Axis::Short operator: Philox blocks run along the rows of S (round 2: tensor cores; round 1: generic SIMT kernel)
        Ss = rb.DenseSkOp(rb.DenseDist(d, m, rb.ScalarDist.Uniform, rb.Axis.Short), rb.RNGState(1997), dt)
        before = rb.counter("tensor_core_launches")
        t = timeit(lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, Ss, 0, 0, A, m, 0.0, B, d))
        print(f"{np.dtype(dt).name} left ColMajor, Axis::Short operator: {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s "
              f"(tensor-core launches {rb.counter('tensor_core_launches') - before})", flush=True)
        Sg = rb.DenseSkOp(rb.DenseDist(d, m, rb.ScalarDist.Gaussian, rb.Axis.Short), rb.RNGState(1997), dt)
        t = timeit(lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, Sg, 0, 0, A, m, 0.0, B, d))
        print(f"{np.dtype(dt).name} left ColMajor, Axis::Short Gaussian operator: {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s", flush=True)
        # the same operators FILLED (S.buff): K-contiguous tiles (Axis::Long) and row-contiguous tiles (Axis::Short: MN-major A
        # operand of the float kernel, row-block copies of the DMMA kernel; tc_xmn = 0 is the generic kernel they used before)
        Sf = rb.DenseSkOp(rb.DenseDist(d, m, rb.ScalarDist.Uniform), rb.RNGState(1997), dt)
        rb.fill_dense(Sf)
        t = timeit(lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, Sf, 0, 0, A, m, 0.0, B, d))
        print(f"{np.dtype(dt).name} left ColMajor, FILLED Axis::Long operator: {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s", flush=True)
        del Sf
        Sfs = rb.DenseSkOp(rb.DenseDist(d, m, rb.ScalarDist.Uniform, rb.Axis.Short), rb.RNGState(1997), dt)
        rb.fill_dense(Sfs)
        before = rb.counter("tensor_core_launches")
        t = timeit(lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, Sfs, 0, 0, A, m, 0.0, B, d))
        print(f"{np.dtype(dt).name} left ColMajor, FILLED Axis::Short operator: {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s "
              f"(tensor-core launches {rb.counter('tensor_core_launches') - before})", flush=True)
        t = timeit(lambda: rb.sketch_general("R", "N", "N", d, n, m, 1.0, Sfs, 0, 0, A, n, 0.0, B, n))
        print(f"{np.dtype(dt).name} left RowMajor, FILLED Axis::Short operator: {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s", flush=True)
        rb.set_option("tc_xmn", 0)
        t = timeit(lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, Sfs, 0, 0, A, m, 0.0, B, d), reps=2)
        print(f"{np.dtype(dt).name} left ColMajor, FILLED Axis::Short operator, generic SIMT kernel (tc_xmn = 0): {t:.3f} ms, "
              f"{flops / t / 1e9:.1f} TFLOP/s", flush=True)
        rb.set_option("tc_xmn", 1)
        del Sfs
        rb.set_option("dense_path", 1)
        t = timeit(lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, Ss, 0, 0, A, m, 0.0, B, d), reps=2)
        print(f"{np.dtype(dt).name} left ColMajor, Axis::Short operator, generic SIMT kernel: {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s", flush=True)
        rb.set_option("dense_path", 3 if dt == np.float32 else 1)
        t = timeit(lambda: rb.sketch_general("R", "N", "N", d, n, m, 1.0, S, 0, 0, A, n, 0.0, B, n), reps=2)
        print(f"{np.dtype(dt).name} left RowMajor, generic SIMT kernel: {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s", flush=True)
        rb.set_option("dense_path", 0)


if __name__ == "__main__":
    main()
