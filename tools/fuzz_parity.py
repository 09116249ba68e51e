#!/usr/bin/env python
"""Randomised differential test (GPU library against the C oracle): random shapes, layouts, transposes, operator windows,
leading dimensions, families, axes, alpha / beta and data types for the dense sketch (left and right), the SASO sketch,
sketch_sparse (CSR / CSC / COO, int32 / int64 indices) and fill_dense windows. Sizes are drawn so that every kernel family is reached (tensor-core float / DMMA double with single
CTAs, CTA pairs, panels; the generic kernels; binned and atomic SASO kernels).

    python tools/fuzz_parity.py [seconds] [seed]          prints one line per failure, a summary at the end; exit code 1 on failure
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
import randblas_b200 as rb  # noqa: E402
from gpu_impl import Gpu  # noqa: E402


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def pick(rng, xs):
    return xs[int(rng.integers(len(xs)))]


def dims(rng, big):
    if big:
        return int(rng.integers(64, 400)), int(rng.integers(40, 700)), int(rng.integers(1000, 9000))
    return int(rng.integers(1, 70)), int(rng.integers(1, 70)), int(rng.integers(1, 300))


def dense_case(rng, gpu, port, ctr, key):
    dt = pick(rng, [np.float32, np.float64])
    tol = 1e-5 if dt == np.float32 else 1e-12
    d, n, m = dims(rng, rng.random() < 0.6)
    layout, opS, opA = pick(rng, "CR"), pick(rng, "NT"), pick(rng, "NT")
    fam, ax = pick(rng, "GU"), pick(rng, "LS")
    left = rng.random() < 0.6
    ro, co = int(rng.integers(0, 9)), int(rng.integers(0, 9))
    alpha, beta = pick(rng, [1.0, -0.5, 2.0]), pick(rng, [0.0, 0.0, 1.0, -1.5])
    # op(S) is d x m (left) or n x d with contraction over n (right: B(m x d) = op(A)(m x n) op(S)(n x d))
    if left:
        Dr, Dc = ((d, m) if opS == "N" else (m, d))
    else:
        Dr, Dc = ((n, d) if opS == "N" else (d, n))
    Dr, Dc = Dr + ro + int(rng.integers(0, 5)), Dc + co + int(rng.integers(0, 5))
    rA, cA = ((m, n) if opA == "N" else (n, m))
    al = 4 if dt == np.float32 else 2
    lda = (rA if layout == "C" else cA) + int(rng.integers(0, 3)) * al
    if rng.random() < 0.7:
        lda = (lda + al - 1) // al * al               # 16-byte aligned columns / rows: what the tensor-core kernels take
    A = rng.standard_normal((cA if layout == "C" else rA) * lda).astype(dt)
    rB, cB = ((d, n) if left else (m, d))
    ldb = (rB if layout == "C" else cB) + int(rng.integers(0, 3))
    B0 = rng.standard_normal((cB if layout == "C" else rB) * ldb).astype(dt)
    B1, B2 = B0.copy(), B0.copy()
    # a third of the operators are FILLED first (S.buff: the materialised-operator kernels, K- and row-contiguous tiles)
    prefill = 1 if rng.random() < 0.35 else 0
    what = ("dense", "left" if left else "right", np.dtype(dt).name, layout, opS, opA, d, n, m, Dr, Dc, fam, ax, ro, co, lda, ldb, alpha, beta,
            "filled" if prefill else "unfilled")
    if left:
        gpu.lskge3(layout, opS, opA, d, n, m, dt(alpha), (Dr, Dc, fam, ax), ctr, key, ro, co, A, lda, dt(beta), B1, ldb, prefill=prefill)
        port.lskge3(layout, opS, opA, d, n, m, dt(alpha), (Dr, Dc, fam, ax), ctr, key, ro, co, A, lda, dt(beta), B2, ldb)
    else:
        gpu.rskge3(layout, opA, opS, m, d, n, dt(alpha), A, lda, (Dr, Dc, fam, ax), ctr, key, ro, co, dt(beta), B1, ldb, prefill=prefill)
        port.rskge3(layout, opA, opS, m, d, n, dt(alpha), A, lda, (Dr, Dc, fam, ax), ctr, key, ro, co, dt(beta), B2, ldb)
    e = relerr(B1, B2)
    return what, e, e < tol


def saso_case(rng, gpu, port, ctr, key):
    dt = pick(rng, [np.float32, np.float64])
    tol = 1e-5 if dt == np.float32 else 1e-12
    big = rng.random() < 0.5
    d = int(rng.integers(300, 2100)) if big else int(rng.integers(2, 60))
    n = int(rng.integers(16, 300)) if big else int(rng.integers(1, 40))
    m = int(rng.integers(5000, 40000)) if big else int(rng.integers(d, 400) if d < 400 else d)
    k = int(rng.integers(1, min(d, 12) + 1))
    layout, opA = pick(rng, "CR"), pick(rng, "NT")
    left = rng.random() < 0.6
    alpha, beta = pick(rng, [1.0, -0.5, 2.0]), pick(rng, [0.0, 0.0, 1.0, -1.5])
    co = int(rng.integers(0, 6))
    # wide SASO d x (m + extra) used as is (left) or transposed on the right
    Dr, Dc = d, m + co + int(rng.integers(0, 4))
    rA, cA = ((m, n) if opA == "N" else (n, m))
    lda = (rA if layout == "C" else cA)
    A = rng.standard_normal(rA * cA).astype(dt)
    if left:
        rB, cB = d, n
    else:
        rB, cB = n, d
    ldb = (rB if layout == "C" else cB) + int(rng.integers(0, 3))
    B0 = rng.standard_normal((cB if layout == "C" else rB) * ldb).astype(dt)
    B1, B2 = B0.copy(), B0.copy()
    what = ("saso", "left" if left else "right", np.dtype(dt).name, layout, opA, d, n, m, k, co, ldb, alpha, beta)
    if left:
        gpu.lskges(layout, "N", opA, d, n, m, dt(alpha), (Dr, Dc, k, "S"), ctr, key, 0, co, A, lda, dt(beta), B1, ldb)
        port.lskges(layout, "N", opA, d, n, m, dt(alpha), (Dr, Dc, k, "S"), ctr, key, 0, co, A, lda, dt(beta), B2, ldb)
    else:
        # B(n x d) = op(A)(n x m) * S^T(m x d): A is stored n x m (opA = N) or m x n (opA = T)
        rA, cA = ((n, m) if opA == "N" else (m, n))
        lda = (rA if layout == "C" else cA)
        gpu.rskges(layout, opA, "T", n, d, m, dt(alpha), A, lda, (Dr, Dc, k, "S"), ctr, key, 0, co, dt(beta), B1, ldb)
        port.rskges(layout, opA, "T", n, d, m, dt(alpha), A, lda, (Dr, Dc, k, "S"), ctr, key, 0, co, dt(beta), B2, ldb)
    e = relerr(B1, B2)
    return what, e, e < tol


def fill_sparse_case(rng, gpu, port, ctr, key):
    # SASO index / sign arrays, bit for bit: thread-per-vector kernel (vec_nnz 2, 4, 8, 16) and lane-per-entry kernel
    short = int(rng.integers(1, 3000))
    long_ = int(rng.integers(short, 20000))
    k = int(pick(rng, [1, 2, 3, 4, 5, 8, 8, 16, 16, 7, 32])) if rng.random() < 0.8 else int(rng.integers(1, 40))
    k = max(1, min(k, short))
    r, c = (short, long_) if rng.random() < 0.5 else (long_, short)
    idt = pick(rng, [np.int32, np.int64])
    vdt = pick(rng, [np.float32, np.float64])
    if rng.random() < 0.3:
        ctr = ol.ctr_add(ctr, (1 << 32) - int(rng.integers(1, 50000)))      # the low counter word wraps inside the run
    a = gpu.fill_sparse(r, c, k, "S", ctr, key, vdt, idt)
    b = port.fill_sparse(r, c, k, "S", ctr, key, vdt, idt)
    ok = all(np.array_equal(x, y) for x, y in zip(a, b))
    return ("fill_sparse", r, c, k, np.dtype(idt).name, np.dtype(vdt).name), 0.0 if ok else 1.0, ok


def fill_case(rng, gpu, port, ctr, key):
    dt = pick(rng, [np.float32, np.float64])
    Dr, Dc = int(rng.integers(1, 300)), int(rng.integers(1, 3000))
    if rng.random() < 0.5:
        Dr, Dc = Dc, Dr
    fam, ax, layout = pick(rng, "GU"), pick(rng, "LS"), pick(rng, "CR")
    nr, nc = int(rng.integers(1, Dr + 1)), int(rng.integers(1, Dc + 1))
    ro, co = int(rng.integers(0, Dr - nr + 1)), int(rng.integers(0, Dc - nc + 1))
    got, n1 = gpu.fill_dense_unpacked(layout, Dr, Dc, fam, ax, nr, nc, ro, co, ctr, key, dt)
    want, n2 = port.fill_dense_unpacked(layout, Dr, Dc, fam, ax, nr, nc, ro, co, ctr, key, dt)
    what = ("fill", np.dtype(dt).name, layout, Dr, Dc, fam, ax, nr, nc, ro, co)
    if fam == "U":
        ok = np.array_equal(got, want)
        e = 0.0 if ok else 1.0
    else:
        g = got.astype(np.float32).view(np.int32).astype(np.int64)
        w = want.astype(np.float32).view(np.int32).astype(np.int64)
        e = float(np.abs(g - w).max())
        ok = e <= 2
    return what, e, ok and list(n1) == list(n2)


def sksp_case(rng, gpu, port, ctr, key):
    import scipy.sparse as sp
    dt = pick(rng, [np.float32, np.float64])
    tol = 1e-5 if dt == np.float32 else 1e-12
    big = rng.random() < 0.4
    d = int(rng.integers(64, 520)) if big else int(rng.integers(1, 50))
    n = int(rng.integers(30, 400)) if big else int(rng.integers(1, 60))
    m = int(rng.integers(500, 5000)) if big else int(rng.integers(1, 200))
    layout, opS, opA = pick(rng, "CR"), pick(rng, "NT"), pick(rng, "NT")
    fam, ax, fmt = pick(rng, "GU"), pick(rng, "LS"), int(rng.integers(0, 3))
    idt = pick(rng, [np.int32, np.int64])
    left = rng.random() < 0.6
    ro, co = int(rng.integers(0, 6)), int(rng.integers(0, 6))
    alpha, beta = pick(rng, [1.0, -0.5]), pick(rng, [0.0, 1.0, -1.5])
    dens = float(pick(rng, [0.002, 0.02, 0.15]))
    if left:       # B(d x n) = op(S)(d x m) op(A)(m x n)
        Dr, Dc = ((d, m) if opS == "N" else (m, d))
        ra, ca = ((m, n) if opA == "N" else (n, m))
        rB, cB = d, n
    else:          # B(m x d) = op(A)(m x n) op(S)(n x d)
        Dr, Dc = ((n, d) if opS == "N" else (d, n))
        ra, ca = ((m, n) if opA == "N" else (n, m))
        rB, cB = m, d
    Dr, Dc = Dr + ro + int(rng.integers(0, 4)), Dc + co + int(rng.integers(0, 4))
    M = sp.random(ra, ca, density=dens, random_state=int(rng.integers(1 << 30)), dtype=np.float64)
    if fmt == 0:
        Mc = M.tocsr(); Mc.sort_indices()
        spA = (ra, ca, Mc.nnz, Mc.data.astype(dt), Mc.indptr.astype(np.int64), Mc.indices.astype(np.int64))
    elif fmt == 1:
        Mc = M.tocsc(); Mc.sort_indices()
        spA = (ra, ca, Mc.nnz, Mc.data.astype(dt), Mc.indices.astype(np.int64), Mc.indptr.astype(np.int64))
    else:
        Mc = M.tocoo()
        spA = (ra, ca, Mc.nnz, Mc.data.astype(dt), Mc.row.astype(np.int64), Mc.col.astype(np.int64))
    ldb = (rB if layout == "C" else cB) + int(rng.integers(0, 3))
    B0 = rng.standard_normal((cB if layout == "C" else rB) * ldb).astype(dt)
    B1, B2 = B0.copy(), B0.copy()
    what = ("sksp", "left" if left else "right", np.dtype(dt).name, fmt, layout, opS, opA, d, n, m, Dr, Dc, fam, ax, ro, co, ldb, alpha, beta, dens)
    if left:
        gpu.lsksp3(fmt, layout, opS, opA, d, n, m, dt(alpha), (Dr, Dc, fam, ax), ctr, key, ro, co, spA, dt(beta), B1, ldb, idx_dtype=idt)
        port.lsksp3(fmt, layout, opS, opA, d, n, m, dt(alpha), (Dr, Dc, fam, ax), ctr, key, ro, co, spA, dt(beta), B2, ldb)
    else:
        gpu.rsksp3(fmt, layout, opA, opS, m, d, n, dt(alpha), spA, (Dr, Dc, fam, ax), ctr, key, ro, co, dt(beta), B1, ldb, idx_dtype=idt)
        port.rsksp3(fmt, layout, opA, opS, m, d, n, dt(alpha), spA, (Dr, Dc, fam, ax), ctr, key, ro, co, dt(beta), B2, ldb)
    e = relerr(B1, B2)
    return what, e, e < tol


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    gpu, port = Gpu(), ol.port()
    port.set_threads(os.cpu_count() or 1)
    t0 = time.time()
    counts, fails, worst = {}, 0, {}
    before = rb.counter("tensor_core_launches")
    while time.time() - t0 < budget:
        ctr, key = ol.state_from_u64(int(rng.integers(0, 2**31)))
        case = pick(rng, [dense_case, dense_case, dense_case, saso_case, saso_case, fill_case, fill_sparse_case, sksp_case])
        try:
            what, e, ok = case(rng, gpu, port, ctr, key)
        except Exception as ex:                       # an error one side raises and the other does not is a finding too
            what, e, ok = (case.__name__, repr(ex)[:200]), float("nan"), False
        counts[what[0]] = counts.get(what[0], 0) + 1
        worst[what[0]] = max(worst.get(what[0], 0.0), e if e == e else 0.0)
        if not ok:
            fails += 1
            print("FAIL", what, e, flush=True)
    print(f"fuzz: {counts}, worst {worst}, tensor-core launches {rb.counter('tensor_core_launches') - before}, failures {fails}", flush=True)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
