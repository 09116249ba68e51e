"""GPU parity tests added in round 2 (run with -m gpu on the B200 box), all through the C ABI:

* the BASELINE.json configurations at their OWN shapes against the compiled reference (oracle/_ref): C1 in full,
  a row block of C3 with a non-zero operator offset, a row slice of a C5 column shard;
* the public spmm entry points and the format conversions against the reference's own left_spmm / right_spmm /
  coo_to_csr / coo_to_csc / csr_to_coo / csc_to_coo (not scipy);
* random_coo bit for bit against the reference's sequential PhiloxStream, random_csr / random_csc structure, and
  the CSR column partition;
* the m-sharded left sketch of the C ABI (rb_lskge3_mshard_*): one rank in-process, two ranks under torchrun with
  CUDA + NCCL (skipped below two devices), and the single-process rb_comm_init form through the C++ drop-in test.
"""
import itertools
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from gpu_impl import Gpu
    return Gpu()


def _gauss(rb, torch, rows, cols, seed, dtype):
    """rows x cols standard-normal data made by the library's own fill_dense (bit-exact to the reference's, so the
    CPU side could regenerate it; here it is simply copied back). Returned flat, in DenseDist's natural layout."""
    buf = torch.empty(rows * cols, dtype=dtype, device="cuda")
    rb.fill_dense(rb.DenseDist(rows, cols), buf, rb.RNGState(seed))
    return buf


# ------------------------------------------------------------------------------ BASELINE shapes vs the reference
def test_c1_full_shape_vs_reference(ref):
    """BASELINE.json configs[0] exactly: sketch_general<float>(ColMajor, N, N, d=1024, n=1024, m=100000, 1,
    DenseSkOp(DenseDist(1024, 100000, Uniform)), A, lda=m, 0, B, ldb=d). Tolerance: 1e-5 relative Frobenius
    (north_star); the reference's own componentwise bound (test/test_matmul_cores/linop_common.hh:260-266) is looser."""
    import torch
    import randblas_b200 as rb
    d, n, m = 1024, 1024, 100000
    A = _gauss(rb, torch, m, n, 99, torch.float32)          # tall => ColMajor natural layout, lda = m
    B = torch.zeros(d * n, dtype=torch.float32, device="cuda")
    S = rb.DenseSkOp(rb.DenseDist(d, m, rb.ScalarDist.Uniform), rb.RNGState(1997), np.float32)
    before = rb.counter("tensor_core_launches")
    rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, B, d)
    torch.cuda.synchronize()
    assert rb.counter("tensor_core_launches") == before + 1
    want = np.zeros(d * n, np.float32)
    ctr, key = ol.state_from_u64(1997)
    ref.set_threads(os.cpu_count() or 1)
    ref.lskge3("C", "N", "N", d, n, m, np.float32(1), (d, m, "U", "L"), ctr, key, 0, 0, A.cpu().numpy(), m,
               np.float32(0), want, d)
    err = relerr(B.cpu().numpy(), want)
    assert err < 1e-5, err
    # the CPU result itself carries float-GEMM rounding; against an fp64 product of the same operands the GPU path
    # must be at least as accurate as the contract as well
    Sm, _ = ref.fill_dense_unpacked("R", d, m, "U", "L", d, m, 0, 0, ctr, key, np.float32)
    exact = (torch.from_numpy(Sm.reshape(d, m)).cuda().double() @ A.view(n, m).t().double()).t().contiguous().view(-1)
    err64 = float((B.double() - exact).norm() / exact.norm())
    assert err64 < 1e-5, err64
    print(f"C1 full: rel err vs reference {err:.2e}, vs fp64 product {err64:.2e}")


def test_c3_block_shape_vs_reference(ref):
    """A row block of BASELINE.json configs[2]: d=4096, n=512, 20000 rows of A taken at row 1,000,000 of the
    m = 4,000,000 problem (co_s = 1000000: what rank 2 of 8 computes first), Gaussian double. 1e-12."""
    import torch
    import randblas_b200 as rb
    d, n, mb, m_total, co = 4096, 512, 20000, 4000000, 1000000
    A = _gauss(rb, torch, mb, n, 99, torch.float64)          # ColMajor, lda = mb
    B = torch.zeros(d * n, dtype=torch.float64, device="cuda")
    S = rb.DenseSkOp(rb.DenseDist(d, m_total, rb.ScalarDist.Gaussian), rb.RNGState(1997), np.float64)
    before = rb.counter("tensor_core_launches")
    rb.sketch_general("C", "N", "N", d, n, mb, 1.0, S, 0, co, A, mb, 0.0, B, d)
    torch.cuda.synchronize()
    assert rb.counter("tensor_core_launches") == before + 1
    want = np.zeros(d * n, np.float64)
    ctr, key = ol.state_from_u64(1997)
    ref.set_threads(os.cpu_count() or 1)
    ref.lskge3("C", "N", "N", d, n, mb, 1.0, (d, m_total, "G", "L"), ctr, key, 0, co, A.cpu().numpy(), mb, 0.0, want, d)
    err = relerr(B.cpu().numpy(), want)
    assert err < 1e-12, err
    print(f"C3 block: rel err vs reference {err:.2e}")


def test_c5_slice_shape_vs_reference(ref):
    """A row slice of a column shard of BASELINE.json configs[4]: d=512, the shard's n_local = 125000 columns, the
    first 100000 rows of the 1e7-row CSR matrix (density 1e-4, made by random_coo -- the reference's generator --
    then CSR), against the reference's sketch_sparse with the matching 512 x 100000 window of the 512 x 1e7 operator."""
    import torch
    import randblas_b200 as rb
    d, n_local, rows, m_total = 512, 125000, 100000, 10000000
    A, _ = rb.random_csr(rows, n_local, 1e-4, rb.RNGState(4242), np.float32, np.int64)
    assert A.ambiguous == 0
    B = torch.zeros(d * n_local, dtype=torch.float32, device="cuda")
    S = rb.DenseSkOp(rb.DenseDist(d, m_total), rb.RNGState(1997), np.float32)
    rb.sketch_sparse("C", "N", "N", d, n_local, rows, 1.0, S, 0, 0, A, 0.0, B, d)
    torch.cuda.synchronize()
    spA = (rows, n_local, A.nnz, A.vals.cpu().numpy(), A.rowptr.cpu().numpy(), A.colidxs.cpu().numpy())
    want = np.zeros(d * n_local, np.float32)
    ctr, key = ol.state_from_u64(1997)
    ref.set_threads(os.cpu_count() or 1)
    ref.lsksp3(0, "C", "N", "N", d, n_local, rows, np.float32(1), (d, m_total, "G", "L"), ctr, key, 0, 0, spA,
               np.float32(0), want, d)
    err = relerr(B.cpu().numpy(), want)
    assert err < 1e-5, err
    print(f"C5 slice: nnz {A.nnz}, rel err vs reference {err:.2e}")


# ------------------------------------------------------------------------------ spmm + conversions vs the reference
def _sp_tuple(M, fmt, dt):
    if fmt == 0:
        M = M.tocsr(); M.sort_indices()
        return (M.shape[0], M.shape[1], M.nnz, M.data.astype(dt), M.indptr.astype(np.int64), M.indices.astype(np.int64))
    if fmt == 1:
        M = M.tocsc(); M.sort_indices()
        return (M.shape[0], M.shape[1], M.nnz, M.data.astype(dt), M.indices.astype(np.int64), M.indptr.astype(np.int64))
    M = M.tocoo()
    return (M.shape[0], M.shape[1], M.nnz, M.data.astype(dt), M.row.astype(np.int64), M.col.astype(np.int64))


def _rb_mat(rb, torch, spA, fmt, idt):
    r, c, nnz, vals, i0, i1 = spA
    cls = (rb.CSRMatrix, rb.CSCMatrix, rb.COOMatrix)[fmt]
    return cls(r, c, nnz, torch.from_numpy(vals).cuda(), torch.from_numpy(i0.astype(idt)).cuda(),
               torch.from_numpy(i1.astype(idt)).cuda())


def test_left_and_right_spmm_vs_reference(ref):
    """rb_spmm_* against RandBLAS::sparse_data::left_spmm / right_spmm themselves (spmm_dispatch.hh:52-219): CSR, CSC,
    COO x both ops of the sparse and of the dense matrix x both layouts x float/double x int32/int64, padded leading
    dimensions, alpha/beta, and a COO submatrix window."""
    import scipy.sparse as sp
    import torch
    import randblas_b200 as rb
    rng = np.random.default_rng(13)
    d, n, m = 37, 23, 211
    for dt, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
        for fmt, idt, lay, opA, opB in itertools.product((0, 1, 2), (np.int32, np.int64), "CR", "NT", "NT"):
            Am = sp.random(*((d, m) if opA == "N" else (m, d)), density=0.08, random_state=int(rng.integers(1 << 30)),
                           format="coo", dtype=np.float64)
            spA = _sp_tuple(Am, fmt, dt)
            # left: C(d x n) = alpha op(A)(d x m) op(B)(m x n) + beta C
            rB, cB = (m, n) if opB == "N" else (n, m)
            ldb = (rB if lay == "C" else cB) + 1
            ldc = (d if lay == "C" else n) + 2
            Bbuf = rng.standard_normal(ldb * (cB if lay == "C" else rB)).astype(dt)
            C0 = rng.standard_normal(ldc * (n if lay == "C" else d)).astype(dt)
            want = C0.copy()
            ref.spmm(1, fmt, lay, opA, opB, d, n, m, dt(0.5), spA, 0, 0, Bbuf, ldb, dt(-1.5), want, ldc)
            Cd = torch.from_numpy(C0.copy()).cuda()
            rb.left_spmm(lay, opA, opB, d, n, m, 0.5, _rb_mat(rb, torch, spA, fmt, idt), 0, 0, torch.from_numpy(Bbuf).cuda(),
                         ldb, -1.5, Cd, ldc)
            assert relerr(Cd.cpu().numpy(), want) < tol, ("left", fmt, idt, lay, opA, opB)
            # right: C(n x d) = alpha op(B)(n x m) op(A_sp)(m x d) + beta C
            Am2 = sp.random(*((m, d) if opA == "N" else (d, m)), density=0.08, random_state=int(rng.integers(1 << 30)),
                            format="coo", dtype=np.float64)
            spA2 = _sp_tuple(Am2, fmt, dt)
            rB2, cB2 = (n, m) if opB == "N" else (m, n)
            ldb2 = (rB2 if lay == "C" else cB2) + 1
            ldc2 = (n if lay == "C" else d) + 1
            Bbuf2 = rng.standard_normal(ldb2 * (cB2 if lay == "C" else rB2)).astype(dt)
            C02 = rng.standard_normal(ldc2 * (d if lay == "C" else n)).astype(dt)
            want2 = C02.copy()
            ref.spmm(0, fmt, lay, opB, opA, n, d, m, dt(2.0), spA2, 0, 0, Bbuf2, ldb2, dt(0.25), want2, ldc2)
            Cd2 = torch.from_numpy(C02.copy()).cuda()
            rb.right_spmm(lay, opB, opA, n, d, m, 2.0, torch.from_numpy(Bbuf2).cuda(), ldb2,
                          _rb_mat(rb, torch, spA2, fmt, idt), 0, 0, 0.25, Cd2, ldc2)
            assert relerr(Cd2.cpu().numpy(), want2) < tol, ("right", fmt, idt, lay, opA, opB)
    # COO submatrix window (the reference accepts offsets for COO only)
    big = sp.random(60, 300, density=0.05, random_state=5, format="coo", dtype=np.float64)
    spA = _sp_tuple(big, 2, np.float64)
    Bbuf = rng.standard_normal(m * n)
    want = np.zeros(d * n)
    ref.spmm(1, 2, "R", "N", "N", d, n, m, 1.0, spA, 3, 7, Bbuf, n, 0.0, want, n)
    Cd = torch.zeros(d * n, dtype=torch.float64, device="cuda")
    rb.left_spmm("R", "N", "N", d, n, m, 1.0, _rb_mat(rb, torch, spA, 2, np.int64), 3, 7, torch.from_numpy(Bbuf).cuda(), n,
                 0.0, Cd, n)
    assert relerr(Cd.cpu().numpy(), want) < 1e-12


def test_sparse_format_conversions_vs_reference(ref):
    """coo_to_csr / coo_to_csc / csr_to_coo / csc_to_coo against the reference's conversions.hh:49-121 on unsorted,
    duplicate-free COO input: identical arrays (indices and values), for both scalar types, device and host arrays."""
    import scipy.sparse as sp
    import torch
    import randblas_b200 as rb
    rng = np.random.default_rng(17)
    for (nr, nc, dens) in ((37, 53, 0.2), (2000, 3000, 0.01), (301, 7, 0.3)):
        M = sp.random(nr, nc, density=dens, random_state=int(rng.integers(1 << 30)), format="coo", dtype=np.float64)
        perm = rng.permutation(M.nnz)
        for idt, dt, on_dev in itertools.product((np.int32, np.int64), (np.float32, np.float64), (True, False)):
            vals, rows, cols = M.data[perm].astype(dt), M.row[perm].astype(np.int64), M.col[perm].astype(np.int64)
            mk = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()) if on_dev else np.ascontiguousarray
            back = (lambda t: t.cpu().numpy()) if on_dev else (lambda t: t)
            coo = rb.COOMatrix(nr, nc, M.nnz, mk(vals), mk(rows.astype(idt)), mk(cols.astype(idt)))
            csr, csc = rb.coo_to_csr(coo), rb.coo_to_csc(coo)
            rv, ri, rp = ref.coo_to_compressed(0, nr, nc, vals, rows, cols)
            assert np.array_equal(back(csr.rowptr), rp) and np.array_equal(back(csr.colidxs), ri)
            assert np.array_equal(back(csr.vals), rv)
            cv, ci, cp = ref.coo_to_compressed(1, nr, nc, vals, rows, cols)
            assert np.array_equal(back(csc.colptr), cp) and np.array_equal(back(csc.rowidxs), ci)
            assert np.array_equal(back(csc.vals), cv)
            assert np.array_equal(back(rb.csr_to_coo(csr).rows), ref.compressed_to_coo(0, nr, nc, rv, ri, rp))
            assert np.array_equal(back(rb.csc_to_coo(csc).cols), ref.compressed_to_coo(1, nr, nc, cv, ci, cp))


# ------------------------------------------------------------------------------ random matrices, column partition
def test_random_coo_bit_exact_vs_reference(ref):
    """rb_random_coo_* against RandBLAS::sparse_data::random_coo (random_matrix.hh:290-355): indices, values and the
    returned state bit for bit, for both scalar and index types, several densities and keys, matrices with no stored
    entry at all, and one large enough to take several scan batches (> 2^22 entries)."""
    import torch
    import randblas_b200 as rb
    cases = [(50, 70, 0.1, 42), (1, 1, 0.5, 0), (7, 3, 0.9, 1), (1000, 999, 1e-3, 7), (300, 400, 1e-7, 3),
             (2000, 3000, 0.02, 1 << 40), (0, 5, 0.3, 2), (5, 9, 0.0, 2), (3000, 4000, 0.45, 11)]
    for (m, n, dens, k) in cases:
        ctr, key = ol.state_from_u64(k)
        ctr = ol.ctr_add(ctr, 0xFFFFFFF0)                     # counter carry into the second limb along the way
        st = rb.RNGState(counter=list(ctr), key=list(key))
        for dt, idt in ((np.float32, np.int64), (np.float64, np.int32)):
            v, r, c, nnz, nxt = ref.random_sparse(2, m, n, dens, ctr, key, dt)
            A, nst = rb.random_coo(m, n, dens, st, dt, idt)
            assert A.ambiguous == 0
            assert A.nnz == nnz, (m, n, dens, k, A.nnz, nnz)
            assert nst.counter == [int(x) for x in nxt], (m, n, dens)
            assert np.array_equal(A.rows.cpu().numpy(), r) and np.array_equal(A.cols.cpu().numpy(), c)
            assert np.array_equal(A.vals.cpu().numpy(), v)
    torch.cuda.synchronize()


def test_random_csr_csc_and_column_block(ref):
    """random_csr / random_csc are the compressed forms of random_coo's matrix (checked through the reference's own
    coo_to_csr / coo_to_csc); csr_column_block equals scipy's column slice, and the column-sharded sketch of the blocks
    equals the corresponding columns of the unsharded sketch_sparse."""
    import scipy.sparse as sp
    import torch
    import randblas_b200 as rb
    m, n, dens = 700, 900, 0.02
    st = rb.RNGState(5)
    ctr, key = ol.state_from_u64(5)
    for dt, idt in ((np.float32, np.int64), (np.float64, np.int32)):
        A, nst = rb.random_csr(m, n, dens, st, dt, idt)
        v, r, c, nnz, nxt = ref.random_sparse(2, m, n, dens, ctr, key, dt)
        rv, ri, rp = ref.coo_to_compressed(0, m, n, v, r, c)
        assert A.nnz == nnz and nst.counter == [int(x) for x in nxt]
        assert np.array_equal(A.rowptr.cpu().numpy(), rp) and np.array_equal(A.colidxs.cpu().numpy(), ri)
        assert np.array_equal(A.vals.cpu().numpy(), rv)
        C, _ = rb.random_csc(m, n, dens, st, dt, idt)
        v2, r2, c2, nnz2, _ = ref.random_sparse(2, n, m, dens, ctr, key, dt)     # the transpose, row-major sorted
        assert C.nnz == nnz2 and np.array_equal(C.rowidxs.cpu().numpy(), c2) and np.array_equal(C.vals.cpu().numpy(), v2)
        ref_ptr = np.searchsorted(r2, np.arange(n + 1))
        assert np.array_equal(C.colptr.cpu().numpy(), ref_ptr)
        # column partition
        M = sp.csr_matrix((rv, ri, rp), shape=(m, n))
        Mc = sp.csc_matrix((C.vals.cpu().numpy(), C.rowidxs.cpu().numpy(), C.colptr.cpu().numpy()), shape=(m, n))
        d = 64
        S = rb.DenseSkOp(rb.DenseDist(d, m), rb.RNGState(1997), dt)
        tdt = torch.float32 if dt == np.float32 else torch.float64
        Bfull = torch.zeros(d * n, dtype=tdt, device="cuda")
        rb.sketch_sparse("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, 0.0, Bfull, d)
        for (c0, c1) in ((0, n), (0, 0), (13, 14), (100, 512), (512, n)):
            blk = rb.csr_column_block(A, c0, c1)
            want = M[:, c0:c1].tocsr(); want.sort_indices()
            assert blk.nnz == want.nnz and blk.n_cols == c1 - c0
            assert np.array_equal(blk.rowptr.cpu().numpy(), want.indptr)
            assert np.array_equal(blk.colidxs.cpu().numpy(), want.indices)
            assert np.array_equal(blk.vals.cpu().numpy(), want.data)
            if c1 > c0:
                Bblk = torch.zeros(d * (c1 - c0), dtype=tdt, device="cuda")
                rb.sketch_sparse("C", "N", "N", d, c1 - c0, m, 1.0, S, 0, 0, blk, 0.0, Bblk, d)
                # same sums in a different order (the kernel reduces into B with floating-point atomics)
                assert relerr(Bblk.cpu().numpy(), Bfull[d * c0: d * c1].cpu().numpy()) < (1e-5 if dt == np.float32 else 1e-12)
            cb = rb.csc_column_block(C, c0, c1)
            wc = Mc[:, c0:c1].tocsc(); wc.sort_indices()
            assert cb.nnz == wc.nnz and np.array_equal(cb.colptr.cpu().numpy(), wc.indptr)
            assert np.array_equal(cb.rowidxs.cpu().numpy(), wc.indices)


# ------------------------------------------------------------------------------ m-sharded sketch in the C ABI
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_mshard_entry_point_one_rank(gpu, ref, dt):
    """rb_lskge3_mshard_* with a communicator of one rank (no NCCL involved): both modes, beta != 0, both layouts,
    a non-trivial operator window, against the reference's sketch_general."""
    import torch
    import randblas_b200 as rb
    from randblas_b200.sharding import Comm, lskge3_mshard
    comm = Comm(1, 0)
    assert comm.info()["nranks"] == 1
    rng = np.random.default_rng(3)
    ctr, key = ol.state_from_u64(1997)
    tol = 1e-5 if dt == np.float32 else 1e-12
    for layout, opS, opA, fam in (("C", "N", "N", "G"), ("R", "N", "N", "U"), ("C", "T", "N", "G"), ("R", "N", "T", "U")):
        d, n, m, ro, co = 48, 20, 1003, 2, 5
        Dr, Dc = (d + 7, m + 9) if opS == "N" else (m + 9, d + 7)
        rA, cA = (m, n) if opA == "N" else (n, m)
        lda = rA if layout == "C" else cA
        ldb = d if layout == "C" else n
        A = rng.standard_normal(rA * cA).astype(dt)
        B0 = rng.standard_normal(d * n).astype(dt)
        S = rb.DenseSkOp(rb.DenseDist(Dr, Dc, fam), rb.RNGState(1997), dt)
        for mode, beta in ((0, 0.0), (1, 0.0), (0, -0.5), (1, 2.0)):
            want = B0.copy()
            ref.lskge3(layout, opS, opA, d, n, m, dt(1.5), (Dr, Dc, fam, "L"), ctr, key, ro, co, A, lda, dt(beta), want, ldb)
            Bd = torch.from_numpy(B0.copy()).cuda()
            lskge3_mshard(comm, layout, opS, opA, d, n, m, 1.5, S, ro, co, torch.from_numpy(A).cuda(), lda, beta, Bd, mode)
            torch.cuda.synchronize()
            assert relerr(Bd.cpu().numpy(), want) < tol, (layout, opS, opA, mode, beta)
    # rb_lskges_mshard_*: the same entry point for an unsampled SASO operator
    from randblas_b200.sharding import lskges_mshard
    for layout, opS, k in (("C", "N", 4), ("R", "N", 8), ("C", "T", 3)):
        d, n, m, ro, co = 48, 20, 1003, 0, 6
        Dr, Dc = (d, m + 9) if opS == "N" else (m + 9, d)
        lda = m if layout == "C" else n
        ldb = d if layout == "C" else n
        A = rng.standard_normal(m * n).astype(dt)
        B0 = rng.standard_normal(d * n).astype(dt)
        S = rb.SparseSkOp(rb.SparseDist(Dr, Dc, k), rb.RNGState(1997), dtype=dt)
        for mode, beta in ((0, 0.0), (1, -0.5)):
            want = B0.copy()
            ref.lskges(layout, opS, "N", d, n, m, dt(1.5), (Dr, Dc, k, "S"), ctr, key, ro if opS == "N" else co,
                       co if opS == "N" else ro, A, lda, dt(beta), want, ldb)
            Bd = torch.from_numpy(B0.copy()).cuda()
            lskges_mshard(comm, layout, opS, "N", d, n, m, 1.5, S, ro if opS == "N" else co, co if opS == "N" else ro,
                          torch.from_numpy(A).cuda(), lda, beta, Bd, mode)
            torch.cuda.synchronize()
            assert relerr(Bd.cpu().numpy(), want) < tol, ("saso", layout, opS, k, mode, beta)
        assert S.nnz < 0
    comm.destroy()


def test_mshard_two_ranks_cuda_nccl_vs_single_gpu_and_reference():
    """The m-sharded left sketch on TWO GPUs, one process per GPU under torchrun, CUDA kernels + NCCL reduce-scatter /
    all-reduce inside librandblas_b200.so, against the single-GPU sketch of the whole A and against the reference
    (1e-12 double, 1e-5 float). The worker is tests/mshard_worker.py; skipped below two devices."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (run with gpurun --gpus 2)")
    port = 29400 + os.getpid() % 500
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "mshard_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MSHARD_OK" in r.stdout


# ------------------------------------------------------------------------------ example callers
def _load_example(name):
    import importlib.util
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "examples", name + ".py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    return ex


def test_example_tls_sparse_skop_matches_reference_pipeline(ref):
    """examples/tls_sparse_skop.py (the reference's examples/total-least-squares/tls_sparse_skop.cc) at a reduced size:
    the sampled SASO operator equals the reference's arrays, the sketched data S*[A|b] equals the reference's
    sketch_general for the same seeds (sampled and unsampled operator), and the sketched TLS solution is close to the
    classical one."""
    import torch
    ex = _load_example("tls_sparse_skop")
    m, n = 3000, 60
    sk = 2 * (n + 1)
    ctr, key = ol.state_from_u64(1997)
    for sample_first in (True, False):
        rel, SAB, S, AB = ex.main(m, n, verbose=False, sample_first=sample_first)
        assert rel < 0.5
        want = np.zeros(sk * (n + 1))
        ref.lskges("C", "N", "N", sk, n + 1, m, 1.0, (sk, m, 8, "S"), ctr, key, 0, 0, AB.cpu().numpy(), m, 0.0, want, sk)
        assert relerr(SAB.cpu().numpy(), want) < 1e-12
        if sample_first:
            v, r, c, nnz, _ = ref.fill_sparse(sk, m, 8, "S", ctr, key, np.float64, np.int64)
            assert S.nnz == nnz and np.array_equal(S.rows.cpu().numpy(), r) and np.array_equal(S.cols.cpu().numpy(), c)
            assert np.array_equal(S.vals.cpu().numpy(), v)
    torch.cuda.synchronize()


def test_example_sparse_low_rank_qb_recovers_planted_vectors():
    """examples/svd_rank1_plus_noise.py (the reference's examples/sparse-low-rank-approx/svd_rank1_plus_noise.cc): the QB
    factors reproduce the sparse matrix's action to the noise level and the leading singular triplet is the planted one."""
    import scipy.sparse as sp
    import torch
    ex = _load_example("svd_rank1_plus_noise")
    m, n, vec_nnz = 4000, 2500, 4
    for p in (1, 2):
        s0, cos_u, cos_v, A, Q, B = ex.main(m, n, vec_nnz, p=p, verbose=False)
        assert abs(s0 - 100.0) < 1e-6 and cos_u > 1 - 1e-10 and cos_v > 1 - 1e-10
        k = max(3, vec_nnz)
        M = sp.coo_matrix((A.vals.cpu().numpy(), (A.rows.cpu().numpy(), A.cols.cpu().numpy())), shape=(m, n)).tocsr()
        Qm, Bm = Q.view(k, m).t().cpu().numpy(), B.view(n, k).t().cpu().numpy()
        assert np.linalg.norm(Qm.T @ Qm - np.eye(k)) < 1e-10
        assert np.linalg.norm(Bm - Qm.T @ M) < 1e-10 * 100            # B = Q^T A through right_spmm
        x = np.random.default_rng(0).standard_normal(n)
        assert np.linalg.norm(M @ x - Qm @ (Bm @ x)) < 1e-3 * np.linalg.norm(M @ x)
    torch.cuda.synchronize()


def test_example_spmm_performance_formats_agree():
    """examples/spmm_performance.py (the reference's examples/simple-kernel-benchmarks/spmm_performance.cc): COO, CSR and CSC
    forms of one random sparse matrix through left_spmm and right_spmm give the same product as densify + GEMM (the example
    asserts 1e-12 between formats itself); every timing is positive."""
    ex = _load_example("spmm_performance")
    for (m, n, d, dens) in ((500, 500, 500, 0.01), (2000, 300, 64, 0.02), (300, 2000, 64, 0.02)):
        out = ex.run_config(m, n, d, dens, trials=2, verbose=False)
        assert out["nnz"] > 0 and all(v > 0 for k, v in out.items() if k != "nnz")


# ------------------------------------------------------------------------------ Axis::Short operators on tensor cores
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_short_axis_operators_on_tensor_cores_vs_oracle(gpu, port, dt):
    """Dense operators whose Philox blocks run along the ROWS of op(S) -- Axis::Short distributions and transposed uses
    of tall Axis::Long ones (the reference's case table, dense_skops.hh:187-199) -- take the tcgen05 / DMMA kernels too
    (round 1 sent them to the SIMT kernel). Ragged tiles in all three dimensions, both families, alpha/beta, windows that
    start on a Philox block boundary (tensor cores) and one that does not (generic kernel), left and right sketches,
    K- and Q-contiguous data; against the oracle, with the launch counter proving which kernel ran."""
    import randblas_b200 as rb
    rng = np.random.default_rng(21)
    ctr, key = ol.state_from_u64(1997)
    tol = 1e-5 if dt == np.float32 else 1e-12
    # (layout, opS, d, n, m, D_rows, D_cols, axis, ro, co, family, alpha, beta, expect tensor cores)
    cases = [("C", "N", 200, 300, 5003, 212, 6000, "S", 4, 6, "G", 0.5, -1.5, True),
             ("C", "N", 130, 70, 2500, 140, 9000, "S", 8, 4001, "U", -2.0, 1.0, True),
             ("R", "N", 256, 130, 1031, 256, 1031, "S", 0, 0, "U", 1.0, 0.0, True),
             ("C", "T", 200, 90, 3000, 3100, 204, "S", 7, 0, "G", 1.0, 0.25, True),      # tall + Short, transposed use
             ("C", "N", 200, 300, 5003, 212, 6000, "S", 3, 6, "G", 0.5, -1.5, False)]    # window off the block boundary
    for (lay, opS, d, n, m, Dr, Dc, ax, ro, co, fam, alpha, beta, tc) in cases:
        pad = 4 if dt == np.float32 else 2
        inner = m if lay == "C" else n
        lda = inner + (pad - inner % pad) % pad
        A = rng.standard_normal((n if lay == "C" else m) * lda).astype(dt)
        ldb = (d if lay == "C" else n) + 1
        B0 = rng.standard_normal((n if lay == "C" else d) * ldb).astype(dt)
        B1, B2 = B0.copy(), B0.copy()
        before = rb.counter("tensor_core_launches")
        gpu.lskge3(lay, opS, "N", d, n, m, dt(alpha), (Dr, Dc, fam, ax), ctr, key, ro, co, A, lda, dt(beta), B1, ldb)
        ran_tc = rb.counter("tensor_core_launches") == before + 1
        assert ran_tc == tc, (lay, opS, d, n, m, ax, ro, co, ran_tc)
        port.lskge3(lay, opS, "N", d, n, m, dt(alpha), (Dr, Dc, fam, ax), ctr, key, ro, co, A, lda, dt(beta), B2, ldb)
        assert relerr(B1, B2) < tol, ((lay, opS, d, n, m, ax, ro, co, fam), relerr(B1, B2))
    # right sketch: B(m x d) = A(m x n) S(n x d) with a tall Axis::Short operator (RowMajor natural layout)
    mm, dd, nn = 1500, 100, 977
    for lay in "CR":
        inner = mm if lay == "C" else nn
        pad = 4 if dt == np.float32 else 2
        lda = inner + (pad - inner % pad) % pad
        A = rng.standard_normal((nn if lay == "C" else mm) * lda).astype(dt)
        ldb = (mm if lay == "C" else dd) + 3
        B0 = rng.standard_normal((dd if lay == "C" else mm) * ldb).astype(dt)
        B1, B2 = B0.copy(), B0.copy()
        before = rb.counter("tensor_core_launches")
        gpu.rskge3(lay, "N", "T", mm, dd, nn, dt(-0.5), A, lda, (120, 1000, "G", "S"), ctr, key, 4, 8, dt(2.0), B1, ldb)
        assert rb.counter("tensor_core_launches") == before + 1, ("right", lay)
        port.rskge3(lay, "N", "T", mm, dd, nn, dt(-0.5), A, lda, (120, 1000, "G", "S"), ctr, key, 4, 8, dt(2.0), B2, ldb)
        assert relerr(B1, B2) < tol, (("right", lay), relerr(B1, B2))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_filled_short_axis_operators_on_tensor_cores_vs_oracle(gpu, port, dt):
    """MATERIALISED operators whose vectors run along the rows of op(S) -- a filled (S.buff) Axis::Short operator, or a
    filled tall Axis::Long one used transposed -- were the last dense case on the SIMT kernel. The float kernel now reads their
    tiles as an MN-major A operand of tcgen05.mma (TMA boxes of 32 rows x 32 k, 32-byte-atom swizzle), the DMMA kernel
    copies them with the thread mapping of its row-block generator. Ragged tiles, both data layouts, CTA-pair shapes
    (an even number of row tiles), windows with an aligned and an unaligned first row; vs the oracle (whose operator entries are the same numbers the fill wrote), with the launch counter proving which kernel ran and the tc_xmn = 0 switch (generic kernel) giving
    the same result within the contract."""
    import randblas_b200 as rb
    rng = np.random.default_rng(23)
    ctr, key = ol.state_from_u64(2024)
    tol = 1e-5 if dt == np.float32 else 1e-12
    f32 = dt == np.float32
    # (layout, opS, d, n, m, D_rows, D_cols, axis, ro, co, family, alpha, beta, tensor cores for float?)
    cases = [("C", "N", 200, 300, 5003, 212, 6000, "S", 4, 6, "G", 0.5, -1.5, True),
             ("C", "N", 256, 520, 2500, 256, 9000, "S", 0, 4001, "U", -2.0, 1.0, True),     # two row tiles: CTA pair
             ("R", "N", 256, 130, 1031, 256, 1031, "S", 0, 0, "U", 1.0, 0.0, True),         # Q-contiguous data as well
             ("C", "T", 200, 90, 3000, 3100, 204, "S", 7, 0, "G", 1.0, 0.25, True),        # tall + Short, transposed use
             ("C", "T", 130, 300, 2100, 2200, 132, "L", 3, 0, "U", 1.0, 0.0, False),         # tall + Long, transposed: K-contiguous
             ("C", "N", 200, 300, 5003, 212, 6000, "S", 3, 6, "G", 0.5, -1.5, False)]       # first row not 16-byte aligned
    for (lay, opS, d, n, m, Dr, Dc, ax, ro, co, fam, alpha, beta, tc32) in cases:
        pad = 4 if f32 else 2
        inner = m if lay == "C" else n
        lda = inner + (pad - inner % pad) % pad
        A = rng.standard_normal((n if lay == "C" else m) * lda).astype(dt)
        ldb = (d if lay == "C" else n) + 1
        B0 = rng.standard_normal((n if lay == "C" else d) * ldb).astype(dt)
        B1, B2, B3 = B0.copy(), B0.copy(), B0.copy()
        before = rb.counter("tensor_core_launches")
        gpu.lskge3(lay, opS, "N", d, n, m, dt(alpha), (Dr, Dc, fam, ax), ctr, key, ro, co, A, lda, dt(beta), B1, ldb, prefill=1)
        ran_tc = rb.counter("tensor_core_launches") == before + 1
        if ax == "S":
            assert ran_tc == (tc32 if f32 else True), (lay, opS, d, n, m, ax, ro, co, ran_tc)
        port.lskge3(lay, opS, "N", d, n, m, dt(alpha), (Dr, Dc, fam, ax), ctr, key, ro, co, A, lda, dt(beta), B2, ldb)
        assert relerr(B1, B2) < tol, ((lay, opS, d, n, m, ax, ro, co, fam), relerr(B1, B2))
        if ax == "S":
            rb.set_option("tc_xmn", 0)
            try:
                before = rb.counter("tensor_core_launches")
                gpu.lskge3(lay, opS, "N", d, n, m, dt(alpha), (Dr, Dc, fam, ax), ctr, key, ro, co, A, lda, dt(beta), B3, ldb, prefill=1)
                assert rb.counter("tensor_core_launches") == before, "tc_xmn = 0 sends these operators to the generic kernel"
            finally:
                rb.set_option("tc_xmn", 1)
            assert relerr(B3, B2) < tol and relerr(B1, B3) < 2 * tol
    # right sketch: B(m x d) = A(m x n) S(n x d) with a FILLED tall Axis::Short operator (RowMajor natural layout)
    mm, dd, nn = 1500, 100, 977
    for lay in "CR":
        inner = mm if lay == "C" else nn
        pad = 4 if f32 else 2
        lda = inner + (pad - inner % pad) % pad
        A = rng.standard_normal((nn if lay == "C" else mm) * lda).astype(dt)
        ldb = (mm if lay == "C" else dd) + 3
        B0 = rng.standard_normal((dd if lay == "C" else mm) * ldb).astype(dt)
        B1, B2 = B0.copy(), B0.copy()
        before = rb.counter("tensor_core_launches")
        gpu.rskge3(lay, "N", "T", mm, dd, nn, dt(-0.5), A, lda, (120, 1000, "G", "S"), ctr, key, 4, 8, dt(2.0), B1, ldb, prefill=1)
        assert rb.counter("tensor_core_launches") == before + 1, ("right", lay)
        port.rskge3(lay, "N", "T", mm, dd, nn, dt(-0.5), A, lda, (120, 1000, "G", "S"), ctr, key, 4, 8, dt(2.0), B2, ldb)
        assert relerr(B1, B2) < tol, (("right", lay), relerr(B1, B2))


def test_double_gaussian_operator_through_panels_vs_fused_and_oracle(gpu, port):
    """Double Gaussian operators with K >= 4096 are generated per K panel into scratch memory and multiplied by the
    materialised-operator DMMA kernel (dmma_materialise = 1, the default). With a 16 MB panel the cases below take four to
    five panels (the last one ragged); the result must agree with the fused kernel (dmma_materialise = 0) and with the
    oracle to 1e-12, for operators with blocks along K and along the rows (Axis::Short), both data layouts, alpha/beta."""
    import randblas_b200 as rb
    rng = np.random.default_rng(5)
    ctr, key = ol.state_from_u64(77)
    # (layout, d, n, m, D_rows, D_cols, axis, ro, co, alpha, beta)
    cases = [("C", 300, 260, 20006, 310, 30000, "L", 3, 1001, 1.0, 0.0),
             ("R", 300, 130, 24001, 300, 24001, "L", 0, 0, -0.5, 2.0),
             ("C", 264, 200, 22222, 280, 25000, "S", 8, 5, 2.0, -1.0)]
    try:
        for (lay, d, n, m, Dr, Dc, ax, ro, co, alpha, beta) in cases:
            inner = m if lay == "C" else n
            lda = inner + inner % 2
            A = rng.standard_normal((n if lay == "C" else m) * lda)
            ldb = (d if lay == "C" else n) + 1
            B0 = rng.standard_normal((n if lay == "C" else d) * ldb)
            out = {}
            for mat, mb in ((1, 16), (0, 16)):
                rb.set_option("dmma_materialise", mat)
                rb.set_option("dmma_panel_mb", mb)
                B = B0.copy()
                before = rb.counter("tensor_core_launches")
                gpu.lskge3(lay, "N", "N", d, n, m, alpha, (Dr, Dc, "G", ax), ctr, key, ro, co, A, lda, beta, B, ldb)
                launches = rb.counter("tensor_core_launches") - before
                assert launches == (1 if mat == 0 else -(-m // ((16 << 20) // 8 // d // 1024 * 1024))), (lay, ax, mat, launches)
                out[mat] = B
            want = B0.copy()
            port.lskge3(lay, "N", "N", d, n, m, alpha, (Dr, Dc, "G", ax), ctr, key, ro, co, A, lda, beta, want, ldb)
            assert relerr(out[1], out[0]) < 1e-13, ((lay, ax), relerr(out[1], out[0]))
            assert relerr(out[1], want) < 1e-12, ((lay, ax), relerr(out[1], want))
    finally:
        rb.set_option("dmma_materialise", 1)
        rb.set_option("dmma_panel_mb", 2048)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_reference_componentwise_bound_dense_and_sparse_operators(gpu, port, dt):
    """The reference's own acceptance test (test/test_matmul_cores/linop_common.hh:198-268): the sketch must equal
    alpha * op(S_sub) * op(A) + beta * B0, computed from the MATERIALISED operator, within the componentwise bound
    E = (|alpha| * k * 2 eps) * |S| |A| + |beta| * eps * |B0| (`:260-266`), element by element. Shapes of
    test_lskge3.cc / test_lskges.cc (tiny: generic kernels) and shapes that reach the tensor-core and binned kernels."""
    eps = float(np.finfo(dt).eps)
    rng = np.random.default_rng(3)
    ctr, key = ol.state_from_u64(0)

    def check(B, S, Aop, alpha, beta, B0, k, what):
        want = alpha * (S.astype(np.float64) @ Aop.astype(np.float64)) + beta * B0.astype(np.float64)
        E = (abs(alpha) * k * 2 * eps) * (np.abs(S).astype(np.float64) @ np.abs(Aop).astype(np.float64)) \
            + abs(beta) * eps * np.abs(B0).astype(np.float64)
        bad = np.abs(B.astype(np.float64) - want) > E
        assert not bad.any(), (what, int(bad.sum()), float(np.abs(B - want).max()), float(E.min()))

    # dense operators: (d, n, m, D_rows, D_cols, family, ro, co, alpha, beta)
    for (d, n, m, Dr, Dc, fam, ro, co, alpha, beta) in [(30, 10, 200, 30, 200, "G", 0, 0, 1.0, 0.0),
                                                        (3, 5, 10, 8, 12, "U", 3, 1, 1.0, 0.0),
                                                        (12, 19, 201, 12, 201, "G", 0, 0, 0.5, -1.0),
                                                        (256, 300, 5000, 260, 6000, "U", 4, 8, 2.0, 0.25),
                                                        (200, 260, 8200, 200, 8200, "G", 0, 0, 1.0, 1.0)]:
        A = rng.standard_normal((m, n)).astype(dt)                          # logical m x n, stored ColMajor below
        B0 = rng.standard_normal((d, n)).astype(dt)
        Sfull, _ = port.fill_dense_unpacked("R", Dr, Dc, fam, "L", d, m, ro, co, ctr, key, dt)
        S = Sfull.reshape(d, m)
        Bc = np.asfortranarray(B0).ravel(order="F").copy()
        gpu.lskge3("C", "N", "N", d, n, m, dt(alpha), (Dr, Dc, fam, "L"), ctr, key, ro, co,
                   np.asfortranarray(A).ravel(order="F"), m, dt(beta), Bc, d)
        check(Bc.reshape(n, d).T, S, A, alpha, beta, B0, m, ("dense", d, n, m, fam))
    # SASO operators (test_lskges.cc shapes and a binned-kernel shape): vec_nnz 4 / 8
    for (d, n, m, k, alpha, beta) in [(7, 5, 20, 3, 1.0, 0.0), (12, 19, 201, 4, -1.5, 0.5), (2048, 64, 30000, 8, 1.0, 0.0)]:
        A = rng.standard_normal((m, n)).astype(dt)
        B0 = rng.standard_normal((d, n)).astype(dt)
        vals, rows, cols, nnz, _ = port.fill_sparse(d, m, k, "S", ctr, key, dt)
        S = np.zeros((d, m), np.float64)
        np.add.at(S, (rows[:nnz], cols[:nnz]), vals[:nnz].astype(np.float64))
        Br = B0.ravel().copy()
        gpu.lskges("R", "N", "N", d, n, m, dt(alpha), (d, m, k, "S"), ctr, key, 0, 0, A.ravel(), n, dt(beta), Br, n)
        check(Br.reshape(d, n), S, A, alpha, beta, B0, m, ("saso", d, n, m, k))


def test_panel_path_falls_back_to_fused_kernel_without_memory():
    """The scratch panel of a Gaussian double operator (2 GB here) is an optimisation: when it cannot be allocated the
    call runs the fused kernel instead of failing. Memory is taken away with a torch allocation; the result must equal
    the fused kernel's (same kernel, bit for bit) and the launch counter must show ONE tensor-core launch, not panels."""
    import torch
    import randblas_b200 as rb
    d, n, m = 4096, 64, 65536
    A = torch.randn(m * n, dtype=torch.float64, device="cuda")
    S = rb.DenseSkOp(rb.DenseDist(d, m), rb.RNGState(11), np.float64)
    rb.set_option("dmma_materialise", 0)
    Bf = torch.zeros(d * n, dtype=torch.float64, device="cuda")
    rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, Bf, d)
    torch.cuda.synchronize()
    rb.set_option("dmma_materialise", 1)
    rb.release_workspace()
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    hog = torch.empty(max(free - (1200 << 20), 0), dtype=torch.uint8, device="cuda")      # leaves ~1.2 GB: no 2 GB panel
    try:
        Bp = torch.zeros(d * n, dtype=torch.float64, device="cuda")
        before = rb.counter("tensor_core_launches")
        rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, Bp, d)
        torch.cuda.synchronize()
        assert rb.counter("tensor_core_launches") == before + 1
        assert torch.equal(Bp, Bf)
    finally:
        del hog
        torch.cuda.empty_cache()
        rb.release_workspace()


# ------------------------------------------------------------------------------ SASO apply, double
def test_saso_binned_kernel_double_vs_oracle(gpu, port):
    """The register-tile SASO apply (saso_binned.cu) instantiated for double (16 columns of C per CTA = the same 128 bytes
    per row slice as 32 float columns): forced on shapes that cross tile boundaries -- two row tiles (d > 1024), ragged
    column slices (n % 16 != 0), ragged last chunk, operator windows, every sub-warp group size, alpha/beta -- against
    the oracle at 1e-12, and at a C4-like slice against an fp64 product of the sampled operator."""
    import torch
    import randblas_b200 as rb
    rng = np.random.default_rng(5)
    ctr, key = ol.state_from_u64(1997)
    dt = np.float64
    try:
        rb.set_option("saso_path", 2)
        before = rb.counter("saso_owner_launches")
        for (d, n, m, vn, ro, co) in ((45, 29, 2111, 3, 2, 5), (1500, 34, 3000, 8, 0, 0), (1100, 50, 1537, 17, 7, 3),
                                      (300, 36, 900, 32, 0, 1), (64, 16, 5000, 1, 1, 0), (2048, 20, 777, 5, 0, 0)):
            for rows_mode, opS in [(rm, o) for rm in (1, 0) for o in "NT"]:      # lane per row (default) / 8-lane groups
                rb.set_option("saso_rows", rows_mode)
                Dr, Dc = (d + ro + 3, m + co + 6) if opS == "N" else (m + ro + 6, d + co + 3)
                lda = n + (n % 2)                                # 16-byte aligned rows for TMA
                A = rng.standard_normal(m * lda)
                ldb = n + 1
                B0 = rng.standard_normal(d * ldb)
                for alpha, beta in ((1.0, 0.0), (0.5, -1.5)):
                    B1, B2 = B0.copy(), B0.copy()
                    gpu.lskges("R", opS, "N", d, n, m, dt(alpha), (Dr, Dc, vn, "S"), ctr, key, ro, co, A, lda, dt(beta), B1, ldb)
                    port.lskges("R", opS, "N", d, n, m, dt(alpha), (Dr, Dc, vn, "S"), ctr, key, ro, co, A, lda, dt(beta), B2, ldb)
                    assert relerr(B1, B2) < 1e-12, ((d, n, m, vn, opS, alpha), relerr(B1, B2))
        assert rb.counter("saso_owner_launches") > before, "the binned kernel did not run for double"
        rb.set_option("saso_path", 0)
        d, n, m, vn = 2048, 128, 100000, 8
        S = rb.SparseSkOp(rb.SparseDist(d, m, vn), rb.RNGState(1997), dtype=dt)
        A = torch.randn(m * n, dtype=torch.float64, device="cuda")
        Bo = torch.full((d * n,), 7.0, dtype=torch.float64, device="cuda")
        before = rb.counter("saso_owner_launches")
        rb.sketch_general("R", "N", "N", d, n, m, 1.0, S, 0, 0, A, n, 0.0, Bo, n)
        assert rb.counter("saso_owner_launches") == before + 1
        rb.fill_sparse(S)
        Sd = torch.sparse_coo_tensor(torch.stack([S.rows, S.cols]), S.vals, (d, m))
        want = torch.sparse.mm(Sd, A.view(m, n)).view(-1)
        assert float(torch.linalg.norm(Bo - want) / torch.linalg.norm(want)) < 1e-12
    finally:
        rb.set_option("saso_path", 0)
        rb.set_option("saso_rows", 1)
