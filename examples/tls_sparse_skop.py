#!/usr/bin/env python
"""Total least squares through a sparse sketch -- the reference's examples/total-least-squares/tls_sparse_skop.cc
against this repository's API (same calls and seeds: DenseDist(m, n) data at RNGState(0), noise at RNGState(1),
SparseSkOp<double>(SparseDist(2(n+1), m, 8, Axis::Short), 1997), fill_sparse, sketch_general).

    python examples/tls_sparse_skop.py [m n]          (default 10000 500, as the reference)

fill_dense, fill_sparse and sketch_general run through librandblas_b200.so; torch provides the device buffers and the
two SVDs (LAPACK gesdd in the reference), which are not part of the sketching path."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402
from tls_dense_skop import init_noisy_data, total_least_squares  # noqa: E402


def main(m=10000, n=500, verbose=True, sample_first=True):
    torch.cuda.set_device(0)
    sk_dim = 2 * (n + 1)
    AB = init_noisy_data(m, n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    # tls_sparse_skop.cc:138-146: a SASO (SJLT / OSNAP) with 8 nonzeros per column
    S = rb.SparseSkOp(rb.SparseDist(sk_dim, m, 8, rb.Axis.Short), rb.RNGState(1997), dtype=np.float64)
    if sample_first:
        rb.fill_sparse(S)            # the reference samples explicitly; without it the kernel regenerates S on the fly
    torch.cuda.synchronize()
    t_sample = time.perf_counter() - t0
    t0 = time.perf_counter()
    SAB = torch.zeros(sk_dim * (n + 1), dtype=torch.float64, device="cuda")
    rb.sketch_general(rb.Layout.ColMajor, rb.Op.NoTrans, rb.Op.NoTrans, sk_dim, n + 1, m, 1.0, S, AB, m, 0.0, SAB, sk_dim)
    torch.cuda.synchronize()
    t_sketch = time.perf_counter() - t0
    t0 = time.perf_counter()
    sketch_x = total_least_squares(SAB.view(n + 1, sk_dim).t())
    torch.cuda.synchronize()
    t_solve = time.perf_counter() - t0
    t0 = time.perf_counter()
    true_x = total_least_squares(AB.view(n + 1, m).t())
    torch.cuda.synchronize()
    t_true = time.perf_counter() - t0
    rel = float(torch.linalg.norm(sketch_x - true_x) / torch.linalg.norm(true_x))
    if verbose:
        print(f"Dimensions of the augmented matrix [A|B]   :  {m} by {n + 1}")
        print(f"Embedding dimension                        :  {sk_dim}")
        print(f"Time to sample S                           :  {t_sample:.4f} seconds")
        print(f"Time to compute SAB = S * AB               :  {t_sketch:.4f} seconds")
        print(f"Time to perform TLS on sketched data       :  {t_solve:.4f} seconds")
        print(f"Time for the classical TLS method          :  {t_true:.4f} seconds")
        print(f"||sketch_x - true_x|| / ||true_x||         :  {rel:.6f}")
    return rel, SAB, S, AB


if __name__ == "__main__":
    if len(sys.argv) == 3:
        main(int(sys.argv[1]), int(sys.argv[2]))
    else:
        main()
