#!/usr/bin/env python
"""Small single-launch drivers for ncu captures (never a bench: numbers printed under a profiler are not results).

    python tools/prof_kernels.py fill_gauss_f64 | fill_unif_f32 | saso_apply | saso_fill | dense_f32 | dense_f64 |
                                 dense_f32_mat | dense_f64_mat | sksp

Each runs the kernel 3 times on a reduced-but-still-larger-than-L2 shape of the matching BASELINE config."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402


def main():
    what = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    torch.cuda.set_device(0)
    for kv in os.environ.get("RB_OPTIONS", "").split(","):      # e.g. RB_OPTIONS=tc_pair=0
        if "=" in kv:
            rb.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    if what == "fill_gauss_f64":
        rows, cols = 256, 1000000
        D = rb.DenseDist(8192, cols, rb.ScalarDist.Gaussian, rb.Axis.Long)
        buf = torch.empty(rows * cols, dtype=torch.float64, device="cuda")
        f = lambda: rb.fill_dense_unpacked("R", D, rows, cols, 0, 0, buf, rb.RNGState(1997))
    elif what == "fill_gauss_f32":
        rows, cols = 512, 1000000
        D = rb.DenseDist(8192, cols, rb.ScalarDist.Gaussian, rb.Axis.Long)
        buf = torch.empty(rows * cols, dtype=torch.float32, device="cuda")
        f = lambda: rb.fill_dense_unpacked("R", D, rows, cols, 0, 0, buf, rb.RNGState(1997))
    elif what == "fill_unif_f32":
        rows, cols = 512, 1000000
        D = rb.DenseDist(8192, cols, rb.ScalarDist.Uniform, rb.Axis.Long)
        buf = torch.empty(rows * cols, dtype=torch.float32, device="cuda")
        f = lambda: rb.fill_dense_unpacked("R", D, rows, cols, 0, 0, buf, rb.RNGState(1997))
    elif what == "saso_apply":
        d, m, n, k = 2048, 1000000, 256, 8
        S = rb.SparseSkOp(rb.SparseDist(d, m, k), rb.RNGState(1997), dtype=np.float32)
        A = torch.randn(m * n, dtype=torch.float32, device="cuda")
        B = torch.zeros(d * n, dtype=torch.float32, device="cuda")
        f = lambda: rb.sketch_general("R", "N", "N", d, n, m, 1.0, S, 0, 0, A, n, 0.0, B, n)
    elif what == "dense_f32":
        d, m, n = 1024, 100000, 1024
        S = rb.DenseSkOp(rb.DenseDist(d, m, rb.ScalarDist.Uniform), rb.RNGState(1997), np.float32)
        A = torch.randn(m * n, dtype=torch.float32, device="cuda")
        B = torch.zeros(d * n, dtype=torch.float32, device="cuda")
        f = lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, B, d)
    elif what == "dense_f32_qmajor":
        # right sketch of ColMajor data (the range-finder call A * S): Q-contiguous data tiles
        d, m, n = 1024, 100000, 1024
        St = rb.DenseSkOp(rb.DenseDist(m, d, rb.ScalarDist.Uniform), rb.RNGState(1997), np.float32)
        A = torch.randn(m * n, dtype=torch.float32, device="cuda")
        B = torch.zeros(d * n, dtype=torch.float32, device="cuda")
        f = lambda: rb.sketch_general("C", "N", "N", n, d, m, 1.0, A, n, St, 0, 0, 0.0, B, n)
    elif what == "dense_f32_gauss":
        d, m, n = 1024, 100000, 1024
        S = rb.DenseSkOp(rb.DenseDist(d, m, rb.ScalarDist.Gaussian), rb.RNGState(1997), np.float32)
        A = torch.randn(m * n, dtype=torch.float32, device="cuda")
        B = torch.zeros(d * n, dtype=torch.float32, device="cuda")
        f = lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, B, d)
    elif what == "dense_f64":
        d, m, n = 4096, 32768, 512
        S = rb.DenseSkOp(rb.DenseDist(d, 4000000, rb.ScalarDist.Gaussian), rb.RNGState(1997), np.float64)
        A = torch.randn(m * n, dtype=torch.float64, device="cuda")
        B = torch.zeros(d * n, dtype=torch.float64, device="cuda")
        f = lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, B, d)
    elif what == "cublas_dgemm":
        # the library kernel the C3 roofline is quoted against, at the profiled C3 slice's shape
        d, m, n = 4096, 32768, 512
        X = torch.randn(d, m, dtype=torch.float64, device="cuda")
        Y = torch.randn(m, n, dtype=torch.float64, device="cuda")
        f = lambda: torch.mm(X, Y)
    elif what in ("dense_f32_mat", "dense_f64_mat"):
        # materialised operator (S.buff filled): the XMAT instantiations of the tensor-core kernels
        dt, tdt = (np.float32, torch.float32) if what == "dense_f32_mat" else (np.float64, torch.float64)
        d, m, n = (1024, 100000, 1024) if dt == np.float32 else (4096, 32768, 512)
        S = rb.DenseSkOp(rb.DenseDist(d, m, rb.ScalarDist.Gaussian), rb.RNGState(1997), dt)
        rb.fill_dense(S)
        A = torch.randn(m * n, dtype=tdt, device="cuda")
        B = torch.zeros(d * n, dtype=tdt, device="cuda")
        f = lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, B, d)
    elif what == "saso_fill":
        d, m, k = 2048, 8000000, 8
        S = rb.SparseSkOp(rb.SparseDist(d, m, k), rb.RNGState(1997), dtype=np.float32)
        f = lambda: rb.fill_sparse(S)
    elif what == "sksp":
        d, m, n = 512, 1000000, 125000
        A, _ = rb.random_csr(m, n, 1e-4, rb.RNGState(4242), np.float32, np.int64)      # a row slice of the C5 shard
        S = rb.DenseSkOp(rb.DenseDist(d, 10000000), rb.RNGState(1997), np.float32)
        B = torch.zeros(d * n, dtype=torch.float32, device="cuda")
        f = lambda: rb.sketch_sparse("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, 0.0, B, d)
    else:
        raise SystemExit("unknown target " + what)
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    print("done", what, "launches", rb.counter("kernel_launches"))


if __name__ == "__main__":
    main()
