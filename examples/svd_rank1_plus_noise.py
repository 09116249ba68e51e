#!/usr/bin/env python
"""Low-rank approximation of a SPARSE matrix by a randomized QB decomposition -- the pipeline of the reference's
examples/sparse-low-rank-approx/svd_rank1_plus_noise.cc (qb_decompose_sparse_matrix :209-246, qb_to_svd :248-272)
against this repository's API:

    A = signal_scale * u v^T (u, v sparse unit vectors with vec_nnz nonzeros) + noise (iid sparse, density 1e-4)
    Q, B = QB(A, k) with p power-iteration passes;  A ~= Q B;  SVD of the k x n factor B gives A ~= U diag(s) V^T.

The sparse noise comes from random_coo (the reference's generator, random_matrix.hh:290-355, here on the device),
the test matrix from fill_dense, and every product with the sparse matrix goes through left_spmm / right_spmm of
librandblas_b200.so. torch provides the QR factorisations and the small SVD (LAPACK geqrf/ungqr/gesdd in the reference).

    python examples/svd_rank1_plus_noise.py [m n vec_nnz]     (default 10000 5000 4)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402


def make_matrix(m, n, vec_nnz, signal_scale=1e2, noise_scale=1e-6, prob_nonzero=1e-4):
    """COO matrix signal + noise (entries on the same position are summed), plus the planted unit vectors."""
    # planted vectors: vec_nnz distinct positions each (repeated_fisher_yates, as make_signal_matrix :127-157), +-1/sqrt(vec_nnz)
    idx_u = torch.zeros(vec_nnz, dtype=torch.int64, device="cuda")
    idx_v = torch.zeros(vec_nnz, dtype=torch.int64, device="cuda")
    st = rb.repeated_fisher_yates(vec_nnz, m, 1, idx_u, rb.RNGState(0))
    rb.repeated_fisher_yates(vec_nnz, n, 1, idx_v, st)
    sgn = torch.tensor([1.0 if i % 2 == 0 else -1.0 for i in range(vec_nnz)], dtype=torch.float64, device="cuda")
    u = torch.zeros(m, dtype=torch.float64, device="cuda")
    v = torch.zeros(n, dtype=torch.float64, device="cuda")
    u[idx_u] = sgn / np.sqrt(vec_nnz)
    v[idx_v] = sgn.flip(0) / np.sqrt(vec_nnz)
    sig_rows = idx_u.repeat(vec_nnz)
    sig_cols = idx_v.repeat_interleave(vec_nnz)
    sig_vals = signal_scale * u[sig_rows] * v[sig_cols]
    noise, _ = rb.random_coo(m, n, prob_nonzero, rb.RNGState(1), np.float64, np.int64)
    rows = torch.cat([noise.rows, sig_rows])
    cols = torch.cat([noise.cols, sig_cols])
    vals = torch.cat([noise_scale * noise.vals, sig_vals])
    # sum_of_coo_matrices (:81-124): merge duplicates
    key = rows * n + cols
    uniq, inv = torch.unique(key, return_inverse=True)
    merged = torch.zeros(uniq.numel(), dtype=torch.float64, device="cuda").index_add_(0, inv, vals)
    A = rb.COOMatrix(m, n, int(uniq.numel()), merged, (uniq // n).contiguous(), (uniq % n).contiguous())
    return A, u, v, noise.nnz


def orth(M, rows, cols):
    """householder_orth (:202-207) on a ColMajor rows x cols buffer."""
    Q, _ = torch.linalg.qr(M.view(cols, rows).t())
    M.copy_(Q.t().contiguous().view(-1))


def qb_decompose_sparse_matrix(A, k, p, state):
    """Q (m x k, ColMajor) and B (k x n, ColMajor) with A ~= Q B (svd_rank1_plus_noise.cc:209-246)."""
    m, n = A.n_rows, A.n_cols
    W1 = torch.zeros(m * k, dtype=torch.float64, device="cuda")       # mat_work1 / Q
    W2 = torch.zeros(n * k, dtype=torch.float64, device="cuda")       # mat_work2
    done = 0
    if p % 2 == 0:
        rb.fill_dense(rb.DenseDist(n, k), W2, state)
    else:
        rb.fill_dense(rb.DenseDist(m, k), W1, state)
        rb.left_spmm("C", "T", "N", n, k, m, 1.0, A, 0, 0, W1, m, 0.0, W2, n)
        done += 1
        orth(W2, n, k)
    while p - done > 0:
        rb.left_spmm("C", "N", "N", m, k, n, 1.0, A, 0, 0, W2, n, 0.0, W1, m)
        orth(W1, m, k)
        rb.left_spmm("C", "T", "N", n, k, m, 1.0, A, 0, 0, W1, m, 0.0, W2, n)
        orth(W2, n, k)
        done += 2
    rb.left_spmm("C", "N", "N", m, k, n, 1.0, A, 0, 0, W2, n, 0.0, W1, m)
    orth(W1, m, k)
    B = torch.zeros(k * n, dtype=torch.float64, device="cuda")
    # B = Q^T A: right_spmm(layout, opA = Trans (dense Q), opB = NoTrans (sparse A), k, n, m, ...)
    rb.right_spmm("C", "T", "N", k, n, m, 1.0, W1, m, A, 0, 0, 0.0, B, k)
    return W1, B


def main(m=10000, n=5000, vec_nnz=4, p=2, verbose=True):
    torch.cuda.set_device(0)
    A, u, v, noise_nnz = make_matrix(m, n, vec_nnz)
    k = max(3, vec_nnz)
    Q, B = qb_decompose_sparse_matrix(A, k, p, rb.RNGState(0))
    # qb_to_svd (:248-272)
    Ub, s, Vt = torch.linalg.svd(B.view(n, k).t(), full_matrices=False)
    U = Q.view(k, m).t() @ Ub
    cos_u = float(torch.abs(U[:, 0] @ u))
    cos_v = float(torch.abs(Vt[0, :] @ v))
    if verbose:
        print(f"matrix {m} x {n}: {A.nnz} stored entries ({noise_nnz} noise + {vec_nnz * vec_nnz} signal)")
        print(f"top singular values of the rank-{k} approximation : {[float(x) for x in s[:3]]}")
        print(f"|<u_top, u>| = {cos_u:.12f}   |<v_top, v>| = {cos_v:.12f}")
    return float(s[0]), cos_u, cos_v, A, Q, B


if __name__ == "__main__":
    if len(sys.argv) == 4:
        main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
    else:
        main()
