#!/usr/bin/env python
"""Total least squares through a dense sketch -- the caller of the hot path that the reference ships as
examples/total-least-squares/tls_dense_skop.cc, written against this repository's API (same calls, same seeds:
DenseDist(m, n) data at RNGState(0), noise at RNGState(1), DenseSkOp(DenseDist(2(n+1), m), 1997)).

    python examples/tls_dense_skop.py [m n]          (default 10000 500, as the reference)

Everything on the sketching path (fill_dense, sketch_general) runs through librandblas_b200.so; torch is used for the
device buffers and for the two SVDs (the reference calls LAPACK gesdd there), which are not part of the path."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402


def init_noisy_data(m, n):
    """[A | A 1 + eps], column-major m x (n + 1) (tls_dense_skop.cc:49-69)."""
    AB = torch.empty(m * (n + 1), dtype=torch.float64, device="cuda")
    eps = torch.empty(m, dtype=torch.float64, device="cuda")
    rb.fill_dense(rb.DenseDist(m, n), AB[: m * n], rb.RNGState(0))
    rb.fill_dense(rb.DenseDist(m, 1), eps, rb.RNGState(1))
    # DenseDist(m, n) is tall with Axis::Long: natural layout ColMajor, so AB[:m*n] is A column by column
    A = AB[: m * n].view(n, m).t()
    AB[m * n:] = A.sum(dim=1) + eps
    return AB


def total_least_squares(M):
    """x with (A + E) x = B + R for the smallest [E, R] (tls_dense_skop.cc:71-90); M is rows x (n + 1)."""
    n = M.shape[1] - 1
    Vt = torch.linalg.svd(M, full_matrices=False).Vh
    return -Vt[n, :n] / Vt[n, n]


def main(m=10000, n=500, verbose=True):
    torch.cuda.set_device(0)
    sk_dim = 2 * (n + 1)
    AB = init_noisy_data(m, n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    S = rb.DenseSkOp(rb.DenseDist(sk_dim, m), rb.RNGState(1997), np.float64)
    SAB = torch.zeros(sk_dim * (n + 1), dtype=torch.float64, device="cuda")
    # the operator is never materialised here: fill_dense(S) is optional, the kernel regenerates S tile by tile
    rb.sketch_general(rb.Layout.ColMajor, rb.Op.NoTrans, rb.Op.NoTrans, sk_dim, n + 1, m, 1.0, S, 0, 0, AB, m, 0.0, SAB,
                      sk_dim)
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
        print(f"Time to compute SAB = S * AB (S generated in the kernel) :  {t_sketch:.4f} seconds")
        print(f"Time to perform TLS on sketched data       :  {t_solve:.4f} seconds")
        print(f"Time for the classical TLS method          :  {t_true:.4f} seconds")
        print(f"||sketch_x - true_x|| / ||true_x||         :  {rel:.6f}")
    return rel, sketch_x, true_x


if __name__ == "__main__":
    if len(sys.argv) == 3:
        main(int(sys.argv[1]), int(sys.argv[2]))
    else:
        main()
