#!/usr/bin/env python
"""Sparse-times-dense multiplication by format -- the reference's examples/simple-kernel-benchmarks/spmm_performance.cc
against this repository's API, on the GPU: for each (m, n, d, density) the same random sparse S (m x n) is held as COO, CSR
and CSC and applied from the left, B(m x d) = S A(n x d), and from the right, B(d x n) = A(d x m) S, through left_spmm /
right_spmm; a densify + GEMM line (torch) is timed beside them, as in the reference. Times are CUDA-event medians.

    python examples/spmm_performance.py                              # the reference's default sweep
    python examples/spmm_performance.py m n d density [num_trials]   # one configuration

The sparse matrix comes from random_coo (bit-identical to the reference's generator); the conversions run on the device."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402


def median_ms(f, trials):
    f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(trials):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def run_config(m, n, d, density, trials, verbose=True):
    coo, _ = rb.random_coo(m, n, density, rb.RNGState(0), np.float64)
    fmts = {"COO": coo, "CSR": rb.coo_to_csr(coo), "CSC": rb.coo_to_csc(coo)}
    dense = torch.zeros(m, n, dtype=torch.float64, device="cuda")
    out = {"nnz": coo.nnz}
    # left: B (m x d) = S (m x n) A (n x d), ColMajor dense operands
    A = torch.randn(n * d, dtype=torch.float64, device="cuda")
    B = torch.zeros(m * d, dtype=torch.float64, device="cuda")
    ref = None
    for name, S in fmts.items():
        out["left " + name] = median_ms(lambda: rb.left_spmm("C", "N", "N", m, d, n, 1.0, S, 0, 0, A, n, 0.0, B, m), trials)
        if ref is None:
            ref = B.clone()
        else:
            assert float((B - ref).norm() / ref.norm()) < 1e-12, ("left", name)

    def densify_gemm_left():
        dense.zero_()
        dense.index_put_((coo.rows, coo.cols), coo.vals)
        return dense @ A.view(d, n).t()
    out["left densify+GEMM"] = median_ms(densify_gemm_left, trials)
    assert float((densify_gemm_left().t().contiguous().view(-1) - ref).norm() / ref.norm()) < 1e-12
    # right: B (d x n) = A (d x m) S (m x n)
    A2 = torch.randn(d * m, dtype=torch.float64, device="cuda")
    B2 = torch.zeros(d * n, dtype=torch.float64, device="cuda")
    ref2 = None
    for name, S in fmts.items():
        out["right " + name] = median_ms(lambda: rb.right_spmm("C", "N", "N", d, n, m, 1.0, A2, d, S, 0, 0, 0.0, B2, d), trials)
        if ref2 is None:
            ref2 = B2.clone()
        else:
            assert float((B2 - ref2).norm() / ref2.norm()) < 1e-12, ("right", name)
    out["right densify+GEMM"] = median_ms(lambda: (dense.zero_(), dense.index_put_((coo.rows, coo.cols), coo.vals),
                                                   A2.view(m, d).t() @ dense), trials)
    if verbose:
        print(f"m={m} n={n} d={d} density={density} nnz={coo.nnz}")
        for side in ("left", "right"):
            best = min((out[f"{side} {f}"], f) for f in fmts)
            for f in list(fmts) + ["densify+GEMM"]:
                print(f"   {side:5s} {f:13s} {out[f'{side} {f}'] * 1e3:9.1f} us")
            print(f"   best {side}: {best[1]} ({best[0] * 1e3:.1f} us)")
    return out


def main(argv):
    torch.cuda.set_device(0)
    print("S is m-by-n (sparse), A and B are dense.\n  Left SpMM:  B(m x d) = S(m x n) * A(n x d)\n"
          "  Right SpMM: B(d x n) = A(d x m) * S(m x n)\n")
    if len(argv) >= 4:
        run_config(int(argv[0]), int(argv[1]), int(argv[2]), float(argv[3]), int(argv[4]) if len(argv) > 4 else 20)
        return
    print("=== SQUARE PROBLEMS (d = m = n) ===")
    for s in (100, 200, 500, 1000, 2000):
        run_config(s, s, s, 0.01, 10)
    print("=== RECTANGULAR PROBLEMS ===")
    for (m, n, d, dens) in ((2000, 2000, 200, 0.01), (5000, 5000, 500, 0.001), (5000, 500, 500, 0.01), (500, 5000, 500, 0.01)):
        run_config(m, n, d, dens, 10)


if __name__ == "__main__":
    main(sys.argv[1:])
