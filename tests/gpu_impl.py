"""Adapter giving the CUDA path (through the C ABI, via randblas_b200.api) the same numpy-level surface as
oracle_lib.Port / oracle_lib.Ref, so one test body can run either side."""
import numpy as np
import torch

import randblas_b200 as rb


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def state(ctr, key):
    return rb.RNGState(counter=[int(x) for x in ctr], key=[int(x) for x in key])


class Gpu:
    kind = "cuda"

    def __init__(self, host_buffers=False):
        # host_buffers=True passes numpy arrays straight to the C ABI (staged inside the call)
        self.host = host_buffers

    def _in(self, a):
        return np.ascontiguousarray(a) if self.host else dev(a)

    def _out(self, t, like):
        if self.host:
            like[...] = t
        else:
            like[...] = t.cpu().numpy()

    def philox(self, ctr, key, n_blocks=1):
        out = torch.zeros(4 * n_blocks, dtype=torch.int32, device="cuda")
        rb.philox_words(state(ctr, key), n_blocks, out)
        return out.cpu().numpy().view(np.uint32)

    def fill_dense_unpacked(self, layout, D_rows, D_cols, family, axis, n_rows, n_cols, ro_s, co_s, ctr, key, dtype,
                            ld=0):
        D = rb.DenseDist(D_rows, D_cols, family, axis)
        inner = n_cols if layout == "R" else n_rows
        outer = n_rows if layout == "R" else n_cols
        ldd = ld if ld else inner
        host = np.full(max(outer * ldd, 1), -777.0, dtype)
        buf = host if self.host else dev(host)
        nxt = rb.fill_dense_unpacked(layout, D, n_rows, n_cols, ro_s, co_s, buf, state(ctr, key), ld)
        res = host if self.host else buf.cpu().numpy()
        if ld:
            return res, np.array(nxt.counter, np.uint32)
        return res[: n_rows * n_cols], np.array(nxt.counter, np.uint32)

    def fill_sparse(self, D_rows, D_cols, vec_nnz, axis, ctr, key, dtype, idx_dtype=np.int64):
        D = rb.SparseDist(D_rows, D_cols, vec_nnz, axis)
        vals = np.zeros(D.full_nnz, dtype)
        rows = np.zeros(D.full_nnz, idx_dtype)
        cols = np.zeros(D.full_nnz, idx_dtype)
        v, r, c = self._in(vals), self._in(rows), self._in(cols)
        nnz, nxt = rb.fill_sparse_unpacked_nosub(D, v, r, c, state(ctr, key))
        self._out(v, vals); self._out(r, rows); self._out(c, cols)
        return vals, rows, cols, nnz, np.array(nxt.counter, np.uint32)

    def repeated_fisher_yates(self, k, n, r, ctr, key, idx_dtype=np.int64):
        s = np.zeros(k * r, idx_dtype)
        t = self._in(s)
        nxt = rb.repeated_fisher_yates(k, n, r, t, state(ctr, key))
        self._out(t, s)
        return s, np.array(nxt.counter, np.uint32)

    # --- index-sampling utilities (RandBLAS/util.hh:459-560) ---
    def sample_indices_iid_uniform(self, n, k, ctr, key, idx_dtype=np.int64, rad_dtype=None):
        s = np.full(max(k, 1), -1, idx_dtype)
        r = None if rad_dtype is None else np.zeros(max(k, 1), rad_dtype)
        ts, tr = self._in(s), (None if r is None else self._in(r))
        nxt = rb.sample_indices_iid_uniform(n, k, ts, state(ctr, key), tr)
        self._out(ts, s)
        if r is not None:
            self._out(tr, r)
        return s[:k], (None if r is None else r[:k]), np.array(nxt.counter, np.uint32)

    def sample_indices_iid(self, n, cdf, k, ctr, key, idx_dtype=np.int64):
        s = np.full(max(k, 1), -1, idx_dtype)
        ts, tc = self._in(s), self._in(np.ascontiguousarray(cdf))
        nxt = rb.sample_indices_iid(n, tc, k, ts, state(ctr, key))
        self._out(ts, s)
        return s[:k], np.array(nxt.counter, np.uint32)

    def weights_to_cdf(self, w, error_if_below=None):
        w = np.ascontiguousarray(w).copy()
        t = self._in(w)
        ok = True
        try:
            rb.weights_to_cdf(len(w), t, error_if_below)
        except rb.RandBLASError:
            ok = False
        self._out(t, w)
        return w, ok

    def _dense_op(self, dist, ctr, key, dtype, prefill):
        S = rb.DenseSkOp(rb.DenseDist(*dist), state(ctr, key), dtype)
        if prefill:
            rb.fill_dense(S)
        return S

    def lskge3(self, layout, opS, opA, d, n, m, alpha, dist, ctr, key, ro_s, co_s, A, lda, beta, B, ldb, prefill=0):
        S = self._dense_op(dist, ctr, key, B.dtype, prefill)
        a, b = self._in(A), self._in(B)
        rb.sketch_general(layout, opS, opA, d, n, m, alpha, S, ro_s, co_s, a, lda, beta, b, ldb)
        self._out(b, B)

    def rskge3(self, layout, opA, opS, m, d, n, alpha, A, lda, dist, ctr, key, ro_s, co_s, beta, B, ldb, prefill=0):
        S = self._dense_op(dist, ctr, key, B.dtype, prefill)
        a, b = self._in(A), self._in(B)
        rb.sketch_general(layout, opA, opS, m, d, n, alpha, a, lda, S, ro_s, co_s, beta, b, ldb)
        self._out(b, B)

    def _sparse_op(self, sdist, ctr, key, dtype, prefill):
        S = rb.SparseSkOp(rb.SparseDist(*sdist), state(ctr, key), dtype=dtype)
        if prefill:
            rb.fill_sparse(S)
        return S

    def lskges(self, layout, opS, opA, d, n, m, alpha, sdist, ctr, key, ro_s, co_s, A, lda, beta, B, ldb, prefill=0):
        S = self._sparse_op(sdist, ctr, key, B.dtype, prefill)
        a, b = self._in(A), self._in(B)
        rb.sketch_general(layout, opS, opA, d, n, m, alpha, S, ro_s, co_s, a, lda, beta, b, ldb)
        self._out(b, B)

    def rskges(self, layout, opA, opS, m, d, n, alpha, A, lda, sdist, ctr, key, ro_s, co_s, beta, B, ldb, prefill=0):
        S = self._sparse_op(sdist, ctr, key, B.dtype, prefill)
        a, b = self._in(A), self._in(B)
        rb.sketch_general(layout, opA, opS, m, d, n, alpha, a, lda, S, ro_s, co_s, beta, b, ldb)
        self._out(b, B)

    def _spmat(self, fmt, spA, idx_dtype):
        A_rows, A_cols, nnz, vals, idx0, idx1 = spA
        cls = (rb.CSRMatrix, rb.CSCMatrix, rb.COOMatrix)[fmt]
        return cls(A_rows, A_cols, nnz, self._in(vals), self._in(idx0.astype(idx_dtype)), self._in(idx1.astype(idx_dtype)))

    def lsksp3(self, fmt, layout, opS, opA, d, n, m, alpha, dist, ctr, key, ro_s, co_s, spA, beta, B, ldb,
               idx_dtype=np.int64):
        S = self._dense_op(dist, ctr, key, B.dtype, 0)
        b = self._in(B)
        rb.sketch_sparse(layout, opS, opA, d, n, m, alpha, S, ro_s, co_s, self._spmat(fmt, spA, idx_dtype), beta, b, ldb)
        self._out(b, B)

    def rsksp3(self, fmt, layout, opA, opS, m, d, n, alpha, spA, dist, ctr, key, ro_s, co_s, beta, B, ldb,
               idx_dtype=np.int64):
        S = self._dense_op(dist, ctr, key, B.dtype, 0)
        b = self._in(B)
        rb.sketch_sparse(layout, opA, opS, m, d, n, alpha, self._spmat(fmt, spA, idx_dtype), S, ro_s, co_s, beta, b, ldb)
        self._out(b, B)
