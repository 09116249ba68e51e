"""Host-side mirror of the reference's public API for the sketching hot path.

Names, argument order, defaults and argument checks follow the reference (file:line cited per item, relative
to the reference root); every function that does work forwards to one C-ABI entry point of
librandblas_b200.so. Buffers may be torch CUDA tensors (used in place; the measured mode) or host arrays
(numpy arrays / CPU torch tensors: staged through device memory inside the C call, as a caller of the CPU
reference would pass them).
"""
import ctypes
import math

import numpy as np

from . import _lib
from ._lib import RandBLASError, call

try:  # torch is plumbing only (device memory + streams); the library itself does not need it
    import torch
except Exception:  # pragma: no cover
    torch = None


class Layout:               # blas::Layout (BLAS++ values)
    ColMajor = "C"
    RowMajor = "R"


class Op:                   # blas::Op
    NoTrans = "N"
    Trans = "T"


class Axis:                 # RandBLAS/base.hh:306-312
    Short = "S"
    Long = "L"


class ScalarDist:           # RandBLAS/dense_skops.hh:215-224
    Gaussian = "G"
    Uniform = "U"


def _require(cond, text):
    # randblas_require, RandBLAS/exceptions.hh:152-162
    if not cond:
        raise RandBLASError(f"({text}) was required, but did not hold")


# ------------------------------------------------------------------------------------------------
class RNGState:
    """RandBLAS/base.hh:64-164. counter: 4 x u32 (128-bit little-endian), key: 2 x u32."""

    def __init__(self, k=0, counter=None, key=None):
        if isinstance(k, RNGState):
            counter, key = k.counter, k.key
            k = 0
        if counter is not None or key is not None:
            self.counter = [int(x) & 0xFFFFFFFF for x in (counter if counter is not None else [0, 0, 0, 0])]
            self.key = [int(x) & 0xFFFFFFFF for x in (key if key is not None else [0, 0])]
        else:
            c = (ctypes.c_uint32 * 4)()
            kk = (ctypes.c_uint32 * 2)()
            _lib.lib().rb_rngstate_from_u64(ctypes.c_uint64(int(k)), c, kk)   # base.hh:116-119
            self.counter, self.key = list(c), list(kk)

    def copy(self):
        return RNGState(counter=self.counter, key=self.key)

    def incr(self, n):
        """A new state whose counter is advanced by n (Random123 array.h `incr`)."""
        c = (ctypes.c_uint32 * 4)(*self.counter)
        _lib.lib().rb_ctr_incr(c, ctypes.c_uint64(int(n)))
        return RNGState(counter=list(c), key=self.key)

    def __eq__(self, o):
        return isinstance(o, RNGState) and self.counter == o.counter and self.key == o.key

    def __repr__(self):
        return f"counter : {{{', '.join(map(str, self.counter))}}}\nkey     : {{{', '.join(map(str, self.key))}}}"

    # ctypes views
    def _c(self):
        return (ctypes.c_uint32 * 4)(*self.counter)

    def _k(self):
        return (ctypes.c_uint32 * 2)(*self.key)


def _addr(x):
    # ctypes objects are handed to _lib.call as objects (passed by reference there) so they stay alive
    return x


class DenseDist:
    """RandBLAS/dense_skops.hh:231-350."""

    def __init__(self, n_rows, n_cols, family=ScalarDist.Gaussian, major_axis=Axis.Long):
        n_rows, n_cols = int(n_rows), int(n_cols)
        _require(n_rows > 0, "n_rows > 0")          # :327
        _require(n_cols > 0, "n_cols > 0")          # :328
        info = (ctypes.c_int64 * 3)()
        iso = ctypes.c_double()
        call("rb_dense_dist_info", "qqccpp", n_rows, n_cols, family, major_axis, _addr(info), _addr(iso))
        self.n_rows, self.n_cols, self.family, self.major_axis = n_rows, n_cols, family, major_axis
        self.dim_major, self.dim_minor = int(info[0]), int(info[1])
        self.natural_layout = chr(info[2])
        self.isometry_scale = iso.value

    def sample(self, seed_state, dtype=np.float32):
        return DenseSkOp(self, seed_state, dtype)

    def _t(self):
        return (self.n_rows, self.n_cols, self.family, self.major_axis)


class SparseDist:
    """RandBLAS/sparse_skops.hh:131-246."""

    def __init__(self, n_rows, n_cols, vec_nnz=4, major_axis=Axis.Short):
        n_rows, n_cols, vec_nnz = int(n_rows), int(n_cols), int(vec_nnz)
        _require(n_rows > 0, "n_rows > 0")
        _require(n_cols > 0, "n_cols > 0")
        _require(vec_nnz > 0, "vec_nnz > 0")
        info = (ctypes.c_int64 * 3)()
        iso = ctypes.c_double()
        call("rb_sparse_dist_info", "qqqcpp", n_rows, n_cols, vec_nnz, major_axis, _addr(info), _addr(iso))
        self.n_rows, self.n_cols, self.vec_nnz, self.major_axis = n_rows, n_cols, vec_nnz, major_axis
        self.dim_major, self.dim_minor, self.full_nnz = int(info[0]), int(info[1]), int(info[2])
        self.isometry_scale = iso.value

    def sample(self, seed_state, dtype=np.float32, index_dtype=np.int64):
        return SparseSkOp(self, seed_state, dtype=dtype, index_dtype=index_dtype)


class DenseSkOp:
    """RandBLAS/dense_skops.hh:357-478. `buff` stays None until fill_dense(S); sketching with an unfilled
    operator regenerates it inside the kernel (the reference materialises a temporary, skge.hh:174-181)."""

    def __init__(self, dist, seed_state, dtype=np.float32):
        self.dist = dist
        self.seed_state = RNGState(seed_state)
        nxt = (ctypes.c_uint32 * 4)()
        call("rb_dense_next_state", "qqccpp", dist.n_rows, dist.n_cols, dist.family, dist.major_axis,
             _addr(self.seed_state._c()), _addr(nxt))
        self.next_state = RNGState(counter=list(nxt), key=self.seed_state.key)    # dense_skops.hh:172-185
        self.n_rows, self.n_cols = dist.n_rows, dist.n_cols
        self.own_memory = True
        self.buff = None
        self.layout = dist.natural_layout
        self.dtype = np.dtype(_np_dtype(dtype))


class SparseSkOp:
    """RandBLAS/sparse_skops.hh:289-450. nnz < 0 until fill_sparse(S)."""

    def __init__(self, dist, seed_state, next_state=None, nnz=-1, vals=None, rows=None, cols=None, dtype=np.float32,
                 index_dtype=np.int64):
        self.dist = dist
        self.seed_state = RNGState(seed_state)
        if next_state is None:
            nxt = (ctypes.c_uint32 * 4)()
            call("rb_sparse_next_state", "qqqcpp", dist.n_rows, dist.n_cols, dist.vec_nnz, dist.major_axis,
                 _addr(self.seed_state._c()), _addr(nxt))
            next_state = RNGState(counter=list(nxt), key=self.seed_state.key)     # sparse_skops.hh:266-283
            self.own_memory = True
        else:
            self.own_memory = False                                             # expert constructor, :413-428
        self.next_state = RNGState(next_state)
        self.n_rows, self.n_cols = dist.n_rows, dist.n_cols
        self.nnz, self.vals, self.rows, self.cols = nnz, vals, rows, cols
        self.dtype = np.dtype(_np_dtype(dtype if vals is None else _dtype_of(vals)))
        self.index_dtype = np.dtype(_np_dtype(index_dtype if rows is None else _dtype_of(rows)))


class _SpMat:
    def __init__(self, n_rows, n_cols, nnz, vals, idx0, idx1, fmt):
        self.n_rows, self.n_cols, self.nnz, self.vals = int(n_rows), int(n_cols), int(nnz), vals
        self._idx0, self._idx1, self._fmt = idx0, idx1, fmt
        self.own_memory = False
        self.index_base = 0


class CSRMatrix(_SpMat):    # RandBLAS/sparse_data/csr_matrix.hh:159-168 (view constructor)
    def __init__(self, n_rows, n_cols, nnz, vals, rowptr, colidxs):
        super().__init__(n_rows, n_cols, nnz, vals, rowptr, colidxs, 0)
        self.rowptr, self.colidxs = rowptr, colidxs


class CSCMatrix(_SpMat):    # RandBLAS/sparse_data/csc_matrix.hh:158-167
    def __init__(self, n_rows, n_cols, nnz, vals, rowidxs, colptr):
        super().__init__(n_rows, n_cols, nnz, vals, rowidxs, colptr, 1)
        self.rowidxs, self.colptr = rowidxs, colptr


class COOMatrix(_SpMat):    # RandBLAS/sparse_data/coo_matrix.hh:177-193
    def __init__(self, n_rows, n_cols, nnz, vals, rows, cols):
        super().__init__(n_rows, n_cols, nnz, vals, rows, cols, 2)
        self.rows, self.cols = rows, cols


# ------------------------------------------------------------------------------------------------
def _np_dtype(dt):
    if torch is not None and isinstance(dt, torch.dtype):
        return {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64}[dt]
    return np.dtype(dt).type


def _dtype_of(x):
    if torch is not None and isinstance(x, torch.Tensor):
        return _np_dtype(x.dtype)
    return x.dtype.type


def _ptr(x):
    """Address of a buffer. Buffers are raw memory with explicit leading dimensions, as in the reference, so a
    strided view (A.T, a slice) would be read as if it were dense: reject it."""
    if x is None:
        return 0
    if torch is not None and isinstance(x, torch.Tensor):
        if not x.is_contiguous():
            raise RandBLASError("buffers must be contiguous tensors (pass layout / leading dimensions, not strided views)")
        return x.data_ptr()
    if not isinstance(x, np.ndarray):
        raise RandBLASError(f"unsupported buffer type {type(x).__name__}: expected a torch tensor or a numpy array")
    if not (x.flags.c_contiguous or x.flags.f_contiguous):
        raise RandBLASError("buffers must be contiguous arrays (pass layout / leading dimensions, not strided views)")
    return x.ctypes.data


def _stream(*bufs):
    """Current stream of the device the CUDA buffers live on (0 = default stream for host buffers). All CUDA
    buffers of one call must live on one device, and that device must be the current one."""
    dev = None
    if torch is not None:
        for b in bufs:
            if isinstance(b, torch.Tensor) and b.is_cuda:
                if dev is None:
                    dev = b.device
                elif b.device != dev:
                    raise RandBLASError(f"buffers live on different devices ({dev} and {b.device})")
        if dev is not None:
            if dev.index is not None and dev.index != torch.cuda.current_device():
                raise RandBLASError(f"buffers live on {dev} but the current device is cuda:{torch.cuda.current_device()}")
            return torch.cuda.current_stream(dev).cuda_stream
    return 0


def _same_dtype(*bufs):
    """The scalar type shared by the data buffers of a call (the reference's template parameter T)."""
    dts = {np.dtype(_dtype_of(b)) for b in bufs if b is not None}
    if len(dts) != 1:
        raise RandBLASError(f"buffers of one call must share one scalar type, got {sorted(str(d) for d in dts)}")
    return dts.pop()


def _sfx(dt):
    dt = np.dtype(dt)
    if dt == np.float32:
        return "f32", "f"
    if dt == np.float64:
        return "f64", "d"
    raise TypeError(f"unsupported scalar type {dt}")


def philox_words(state, n_blocks, out):
    """out[4i:4i+4] = Philox4x32_10(state.counter + i, state.key). out: uint32 buffer viewed as int32 if torch."""
    call("rb_philox_words", "ppqpp", _addr(state._c()), _addr(state._k()), int(n_blocks), _ptr(out), _stream(out))
    return out


def boxmuller_words(w0, w1, g0, g1):
    """(g0, g1) = r123::boxmuller(w0, w1) element-wise (RandBLAS/random_gen.hh:62-74); uint32 words as int32 if torch."""
    n = int(w0.numel()) if hasattr(w0, "numel") else int(w0.size)
    call("rb_boxmuller_words", "qppppp", n, _ptr(w0), _ptr(w1), _ptr(g0), _ptr(g1), _stream(g0))
    return g0, g1


def fill_dense_unpacked(layout, D, n_rows, n_cols, ro_s, co_s, buff, seed, ld=0):
    """RandBLAS/dense_skops.hh:563-606. Returns the advanced RNGState."""
    sfx, _ = _sfx(_dtype_of(buff))
    nxt = (ctypes.c_uint32 * 4)()
    call(f"rb_fill_dense_{sfx}", "cqqccqqqqpqpppp", layout, D.n_rows, D.n_cols, D.family, D.major_axis, int(n_rows),
         int(n_cols), int(ro_s), int(co_s), _ptr(buff), int(ld), _addr(seed._c()), _addr(seed._k()), _addr(nxt),
         _stream(buff))
    return RNGState(counter=list(nxt), key=seed.key)


def fill_dense(*args):
    """fill_dense(D, buff, seed) -> RNGState (dense_skops.hh:623-626), or fill_dense(S) (:649-658)."""
    if len(args) == 3:
        D, buff, seed = args
        return fill_dense_unpacked(D.natural_layout, D, D.n_rows, D.n_cols, 0, 0, buff, seed)
    (S,) = args
    if S.own_memory and S.buff is None:
        # the reference allocates host memory with new[]; a GPU library keeps the operator in device memory
        _require(torch is not None, "torch available to allocate S.buff")
        S.buff = torch.empty(S.n_rows * S.n_cols, dtype={np.float32: torch.float32, np.float64: torch.float64}[S.dtype.type],
                             device="cuda")
    _require(S.buff is not None, "S.buff != nullptr")
    fill_dense_unpacked(S.layout, S.dist, S.n_rows, S.n_cols, 0, 0, S.buff, S.seed_state)


def fill_sparse_unpacked_nosub(D, vals, rows, cols, seed_state):
    """RandBLAS/sparse_skops.hh:515-565 (SASO: :526-533, LASO: :534-564). Returns (nnz, next_state)."""
    nnz = ctypes.c_int64(-1)
    nxt = (ctypes.c_uint32 * 4)()
    fn = "rb_fill_sparse_saso" if D.major_axis == Axis.Short else "rb_fill_sparse_laso"
    call(fn, "qqqpppippippp", D.n_rows, D.n_cols, D.vec_nnz, _addr(seed_state._c()),
         _addr(seed_state._k()), _ptr(vals), np.dtype(_dtype_of(vals)).itemsize, _ptr(rows), _ptr(cols),
         np.dtype(_dtype_of(rows)).itemsize, _addr(nnz), _addr(nxt), _stream(vals, rows, cols))
    return nnz.value, RNGState(counter=list(nxt), key=seed_state.key)


def fill_sparse(S):
    """RandBLAS/sparse_skops.hh:587-603."""
    full = S.dist.full_nnz
    if S.own_memory:
        _require(torch is not None, "torch available to allocate the COO arrays")
        tdt = {np.float32: torch.float32, np.float64: torch.float64}[S.dtype.type]
        idt = {np.int32: torch.int32, np.int64: torch.int64}[S.index_dtype.type]
        if S.rows is None:
            S.rows = torch.empty(full, dtype=idt, device="cuda")
        if S.cols is None:
            S.cols = torch.empty(full, dtype=idt, device="cuda")
        if S.vals is None:
            S.vals = torch.empty(full, dtype=tdt, device="cuda")
    _require(S.rows is not None, "S.rows != nullptr")
    _require(S.cols is not None, "S.cols != nullptr")
    _require(S.vals is not None, "S.vals != nullptr")
    S.nnz, _ = fill_sparse_unpacked_nosub(S.dist, S.vals, S.rows, S.cols, S.seed_state)


def repeated_fisher_yates(k, n, r, samples, state):
    """RandBLAS/sparse_skops.hh:259-264. Returns the advanced RNGState."""
    nxt = (ctypes.c_uint32 * 4)()
    call("rb_repeated_fisher_yates", "qqqpipppp", int(k), int(n), int(r), _ptr(samples),
         np.dtype(_dtype_of(samples)).itemsize, _addr(state._c()), _addr(state._k()), _addr(nxt), _stream(samples))
    return RNGState(counter=list(nxt), key=state.key)


def sample_indices_iid_uniform(n, k, samples, state, rademachers=None):
    """RandBLAS/util.hh:515-560: `sample_indices_iid_uniform(n, k, samples, state)` and, with `rademachers`, the
    `<T, sint_t, WriteRademachers = true>` form. Returns the advanced RNGState."""
    nxt = (ctypes.c_uint32 * 4)()
    vb = np.dtype(_dtype_of(rademachers)).itemsize if rademachers is not None else 4
    call("rb_sample_indices_iid_uniform", "qqpipipppp", int(n), int(k), _ptr(samples),
         np.dtype(_dtype_of(samples)).itemsize, _ptr(rademachers), vb, _addr(state._c()), _addr(state._k()), _addr(nxt),
         _stream(samples, rademachers))
    return RNGState(counter=list(nxt), key=state.key)


def sample_indices_iid(n, cdf, k, samples, state):
    """RandBLAS/util.hh:490-513. Returns the advanced RNGState."""
    nxt = (ctypes.c_uint32 * 4)()
    call("rb_sample_indices_iid", "qpiqpipppp", int(n), _ptr(cdf), np.dtype(_dtype_of(cdf)).itemsize, int(k),
         _ptr(samples), np.dtype(_dtype_of(samples)).itemsize, _addr(state._c()), _addr(state._k()), _addr(nxt),
         _stream(samples, cdf))
    return RNGState(counter=list(nxt), key=state.key)


def weights_to_cdf(n, w, error_if_below=None):
    """RandBLAS/util.hh:459-473 (default error_if_below = -sqrt(epsilon<T>), :440-443). In place."""
    dt = np.dtype(_dtype_of(w))
    if error_if_below is None:
        error_if_below = -float(np.sqrt(np.finfo(dt).eps).astype(dt))
    sfx, t = _sfx(dt)
    call(f"rb_weights_to_cdf_{sfx}", "qp" + t + "p", int(n), _ptr(w), float(error_if_below), _stream(w))


def _is_op(x):
    return isinstance(x, (DenseSkOp, SparseSkOp))


def sketch_general(layout, op1, op2, x, y, z, alpha, *rest):
    """The eight overloads of RandBLAS::sketch_general (RandBLAS/skge.hh:756-821, 928-992, 1073-1097, 1175-1199).

    left : sketch_general(layout, opS, opA, d, n, m, alpha, S, [ro_s, co_s,] A, lda, beta, B, ldb)
    right: sketch_general(layout, opA, opS, m, d, n, alpha, A, lda, S, [ro_s, co_s,] beta, B, ldb)
    """
    if _is_op(rest[0]):
        S = rest[0]
        opS, opA, d, n, m = op1, op2, int(x), int(y), int(z)
        if len(rest) == 6:      # full-operator overload: dimension checks of skge.hh:1089-1095
            A, lda, beta, B, ldb = rest[1:]
            if opS == Op.NoTrans:
                _require(S.n_rows == d, "S.n_rows == d"); _require(S.n_cols == m, "S.n_cols == m")
            else:
                _require(S.n_rows == m, "S.n_rows == m"); _require(S.n_cols == d, "S.n_cols == d")
            ro_s = co_s = 0
        else:
            ro_s, co_s, A, lda, beta, B, ldb = rest[1:]
        return _skge(True, layout, opS, opA, d, n, m, alpha, S, int(ro_s), int(co_s), A, int(lda), beta, B, int(ldb))
    A, lda, S = rest[0], rest[1], rest[2]
    opA, opS, m, d, n = op1, op2, int(x), int(y), int(z)
    if len(rest) == 6:          # skge.hh:1191-1197
        beta, B, ldb = rest[3:]
        if opS == Op.NoTrans:
            _require(S.n_rows == n, "S.n_rows == n"); _require(S.n_cols == d, "S.n_cols == d")
        else:
            _require(S.n_rows == d, "S.n_rows == d"); _require(S.n_cols == n, "S.n_cols == n")
        ro_s = co_s = 0
    else:
        ro_s, co_s, beta, B, ldb = rest[3:]
    return _skge(False, layout, opS, opA, d, n, m, alpha, S, int(ro_s), int(co_s), A, int(lda), beta, B, int(ldb))


def _skge(left, layout, opS, opA, d, n, m, alpha, S, ro_s, co_s, A, lda, beta, B, ldb):
    sfx, t = _sfx(_same_dtype(A, B, getattr(S, "buff", None), getattr(S, "vals", None) if getattr(S, "nnz", -1) >= 0 else None))
    st = _stream(A, B, getattr(S, "buff", None))
    seed = S.seed_state
    if isinstance(S, DenseSkOp):
        D = S.dist
        if left:   # dense::lskge3, skge.hh:154-202
            call(f"rb_lskge3_{sfx}", "cccqqq" + t + "qqccppp" + "qqpq" + t + "pqp", layout, opS, opA, d, n, m, alpha,
                 D.n_rows, D.n_cols, D.family, D.major_axis, _addr(seed._c()), _addr(seed._k()), _ptr(S.buff), ro_s,
                 co_s, _ptr(A), lda, beta, _ptr(B), ldb, st)
        else:      # dense::rskge3, skge.hh:307-355
            call(f"rb_rskge3_{sfx}", "cccqqq" + t + "pq" + "qqccppp" + "qq" + t + "pqp", layout, opA, opS, m, d, n,
                 alpha, _ptr(A), lda, D.n_rows, D.n_cols, D.family, D.major_axis, _addr(seed._c()), _addr(seed._k()),
                 _ptr(S.buff), ro_s, co_s, beta, _ptr(B), ldb, st)
        return
    D = S.dist
    if S.nnz >= 0:
        # already sampled (S.vals/rows/cols): coo_view_of_skop + left_spmm/right_spmm, skge.hh:489-490, 621-625
        call(f"rb_coo_apply_{sfx}", "icccqqq" + t + "qqqpppi" + "qqpq" + t + "pqp", 1 if left else 0, layout, opS, opA,
             d, n, m, alpha, S.n_rows, S.n_cols, S.nnz, _ptr(S.vals), _ptr(S.rows), _ptr(S.cols),
             S.index_dtype.itemsize, ro_s, co_s, _ptr(A), lda, beta, _ptr(B), ldb, st)
        return
    if D.major_axis != Axis.Short:
        # LASO, not yet sampled: sample a temporary and drop it (the reference does the same, skge.hh:483-488)
        tmp = SparseSkOp(D, seed, dtype=S.dtype, index_dtype=S.index_dtype)
        fill_sparse(tmp)
        return _skge(left, layout, opS, opA, d, n, m, alpha, tmp, ro_s, co_s, A, lda, beta, B, ldb)
    if left:       # sparse::lskges, skge.hh:465-492
        call(f"rb_lskges_{sfx}", "cccqqq" + t + "qqqpp" + "qqpq" + t + "pqp", layout, opS, opA, d, n, m, alpha,
             D.n_rows, D.n_cols, D.vec_nnz, _addr(seed._c()), _addr(seed._k()), ro_s, co_s, _ptr(A), lda, beta,
             _ptr(B), ldb, st)
    else:          # sparse::rskges, skge.hh:598-626
        call(f"rb_rskges_{sfx}", "cccqqq" + t + "pq" + "qqqpp" + "qq" + t + "pqp", layout, opA, opS, m, d, n, alpha,
             _ptr(A), lda, D.n_rows, D.n_cols, D.vec_nnz, _addr(seed._c()), _addr(seed._k()), ro_s, co_s, beta,
             _ptr(B), ldb, st)


def sketch_symmetric(layout, x, y, alpha, *rest, sym_check_tol=0):
    """RandBLAS/sksy.hh:159-176 (A on the left: B(n x d) = alpha A S + beta B) and :294-312 (S on the left:
    B(d x n) = alpha S A + beta B); A is an n x n symmetric matrix stored as a general one.

    right: sketch_symmetric(layout, n, d, alpha, A, lda, S, ro_s, co_s, beta, B, ldb)
    left : sketch_symmetric(layout, d, n, alpha, S, ro_s, co_s, A, lda, beta, B, ldb)
    """
    if _is_op(rest[0]):
        S, ro_s, co_s, A, lda, beta, B, ldb = rest
        d, n = int(x), int(y)
    else:
        A, lda, S, ro_s, co_s, beta, B, ldb = rest
        n, d = int(x), int(y)
    sfx, t = _sfx(_dtype_of(B))
    call(f"rb_require_symmetric_{sfx}", "cpqq" + t + "p", layout, _ptr(A), n, int(lda), sym_check_tol, _stream(A, B))
    if _is_op(rest[0]):
        return sketch_general(layout, Op.NoTrans, Op.NoTrans, d, n, n, alpha, S, ro_s, co_s, A, lda, beta, B, ldb)
    return sketch_general(layout, Op.NoTrans, Op.NoTrans, n, d, n, alpha, A, lda, S, ro_s, co_s, beta, B, ldb)


def sketch_vector(opS, *args):
    """RandBLAS/skve.hh:141-164 (submatrix) and :233-246 (full operator).

    sketch_vector(opS, d, m, alpha, S, ro_s, co_s, x, incx, beta, y, incy)
    sketch_vector(opS, alpha, S, x, incx, beta, y, incy)
    """
    if len(args) == 7:
        alpha, S, x, incx, beta, y, incy = args
        d, m, ro_s, co_s = S.dist.n_rows, S.dist.n_cols, 0, 0
    else:
        d, m, alpha, S, ro_s, co_s, x, incx, beta, y, incy = args
    _d, _m = (m, d) if opS == Op.Trans else (d, m)
    return sketch_general(Layout.RowMajor, opS, Op.NoTrans, _d, 1, _m, alpha, S, ro_s, co_s, x, incx, beta, y, incy)


def _like(x, n, dtype=None):
    """A new array of n entries living where x lives (torch CUDA tensor or numpy array)."""
    if torch is not None and isinstance(x, torch.Tensor):
        return torch.empty(int(n), dtype=dtype if dtype is not None else x.dtype, device=x.device)
    return np.empty(int(n), dtype=dtype if dtype is not None else x.dtype)


def coo_to_csr(coo):
    """RandBLAS/sparse_data/conversions.hh:101-121. Returns a new CSRMatrix (entries ordered by (row, col))."""
    _require(coo.index_base == 0, "coo.index_base == IndexBase::Zero")
    vals, idx, ptr = _like(coo.vals, coo.nnz), _like(coo.rows, coo.nnz), _like(coo.rows, coo.n_rows + 1)
    call("rb_coo_to_compressed", "iqqqpippipppp", 0, coo.n_rows, coo.n_cols, coo.nnz, _ptr(coo.vals),
         np.dtype(_dtype_of(coo.vals)).itemsize, _ptr(coo.rows), _ptr(coo.cols), np.dtype(_dtype_of(coo.rows)).itemsize,
         _ptr(vals), _ptr(idx), _ptr(ptr), _stream(coo.vals, vals))
    return CSRMatrix(coo.n_rows, coo.n_cols, coo.nnz, vals, ptr, idx)


def coo_to_csc(coo):
    """RandBLAS/sparse_data/conversions.hh:79-99. Returns a new CSCMatrix (entries ordered by (col, row))."""
    _require(coo.index_base == 0, "coo.index_base == IndexBase::Zero")
    vals, idx, ptr = _like(coo.vals, coo.nnz), _like(coo.rows, coo.nnz), _like(coo.rows, coo.n_cols + 1)
    call("rb_coo_to_compressed", "iqqqpippipppp", 1, coo.n_rows, coo.n_cols, coo.nnz, _ptr(coo.vals),
         np.dtype(_dtype_of(coo.vals)).itemsize, _ptr(coo.rows), _ptr(coo.cols), np.dtype(_dtype_of(coo.rows)).itemsize,
         _ptr(vals), _ptr(idx), _ptr(ptr), _stream(coo.vals, vals))
    return CSCMatrix(coo.n_rows, coo.n_cols, coo.nnz, vals, idx, ptr)


def csr_to_coo(csr):
    """RandBLAS/sparse_data/conversions.hh:63-75. vals and cols are shared with the CSR matrix, rows is new."""
    rows = _like(csr.colidxs, csr.nnz)
    call("rb_expand_ptr", "qpqpip", csr.n_rows, _ptr(csr.rowptr), csr.nnz, _ptr(rows),
         np.dtype(_dtype_of(csr.rowptr)).itemsize, _stream(csr.rowptr, rows))
    return COOMatrix(csr.n_rows, csr.n_cols, csr.nnz, csr.vals, rows, csr.colidxs)


def csc_to_coo(csc):
    """RandBLAS/sparse_data/conversions.hh:49-61. vals and rows are shared with the CSC matrix, cols is new."""
    cols = _like(csc.rowidxs, csc.nnz)
    call("rb_expand_ptr", "qpqpip", csc.n_cols, _ptr(csc.colptr), csc.nnz, _ptr(cols),
         np.dtype(_dtype_of(csc.colptr)).itemsize, _stream(csc.colptr, cols))
    return COOMatrix(csc.n_rows, csc.n_cols, csc.nnz, csc.vals, csc.rowidxs, cols)


def left_spmm(layout, opA, opB, d, n, m, alpha, A, ro_a, co_a, B, ldb, beta, C, ldc):
    """RandBLAS::sparse_data::left_spmm (RandBLAS/sparse_data/spmm_dispatch.hh:52-178):
    C(d x n) = alpha * op(A_sp[ro_a:, co_a:])(d x m) * op(B)(m x n) + beta * C."""
    _require(A.index_base == 0, "A.index_base == IndexBase::Zero")      # spmm_dispatch.hh:92
    sfx, t = _sfx(_same_dtype(A.vals, B, C))
    ib = np.dtype(_dtype_of(A._idx0)).itemsize
    call(f"rb_spmm_{sfx}", "iicccqqq" + t + "qqqpppi" + "qqpq" + t + "pqp", 1, A._fmt, layout, opA, opB, int(d), int(n),
         int(m), alpha, A.n_rows, A.n_cols, A.nnz, _ptr(A.vals), _ptr(A._idx0), _ptr(A._idx1), ib, int(ro_a), int(co_a),
         _ptr(B), int(ldb), beta, _ptr(C), int(ldc), _stream(B, C))


def right_spmm(layout, opA, opB, m, d, n, alpha, A, lda, B, i_off, j_off, beta, C, ldc):
    """RandBLAS::sparse_data::right_spmm (spmm_dispatch.hh:180-219):
    C(m x d) = alpha * op(A)(m x n) * op(B_sp[i_off:, j_off:])(n x d) + beta * C, A dense, B sparse."""
    _require(B.index_base == 0, "B.index_base == IndexBase::Zero")
    sfx, t = _sfx(_same_dtype(B.vals, A, C))
    ib = np.dtype(_dtype_of(B._idx0)).itemsize
    call(f"rb_spmm_{sfx}", "iicccqqq" + t + "qqqpppi" + "qqpq" + t + "pqp", 0, B._fmt, layout, opB, opA, int(d), int(n),
         int(m), alpha, B.n_rows, B.n_cols, B.nnz, _ptr(B.vals), _ptr(B._idx0), _ptr(B._idx1), ib, int(i_off), int(j_off),
         _ptr(A), int(lda), beta, _ptr(C), int(ldc), _stream(A, C))


def sketch_sparse(layout, op1, op2, x, y, z, alpha, *rest):
    """RandBLAS/sparse_data/sksp.hh:418-437 (left) and :520-539 (right).

    left : sketch_sparse(layout, opS, opA, d, n, m, alpha, S, ro_s, co_s, A_sp, beta, B, ldb)
    right: sketch_sparse(layout, opA, opS, m, d, n, alpha, A_sp, S, ro_s, co_s, beta, B, ldb)
    """
    if _is_op(rest[0]):
        S, ro_s, co_s, A, beta, B, ldb = rest
        left, opS, opA, d, n, m = True, op1, op2, int(x), int(y), int(z)
    else:
        A, S, ro_s, co_s, beta, B, ldb = rest
        left, opA, opS, m, d, n = False, op1, op2, int(x), int(y), int(z)
    _require(isinstance(S, DenseSkOp), "S is a DenseSkOp")
    _require(A.index_base == 0, "A.index_base == IndexBase::Zero")      # spmm_dispatch.hh:92
    sfx, t = _sfx(_same_dtype(A.vals, B))
    D, seed = S.dist, S.seed_state
    st = _stream(B, A.vals, A._idx0, A._idx1)
    ib = np.dtype(_dtype_of(A._idx0)).itemsize
    if left:
        call(f"rb_lsksp3_{sfx}", "icccqqq" + t + "qqccpp" + "qq" + "qqqpppi" + "qq" + t + "pqp", A._fmt, layout, opS,
             opA, d, n, m, alpha, D.n_rows, D.n_cols, D.family, D.major_axis, _addr(seed._c()), _addr(seed._k()),
             int(ro_s), int(co_s), A.n_rows, A.n_cols, A.nnz, _ptr(A.vals), _ptr(A._idx0), _ptr(A._idx1), ib, 0, 0,
             beta, _ptr(B), int(ldb), st)
    else:
        call(f"rb_rsksp3_{sfx}", "icccqqq" + t + "qqqpppi" + "qq" + "qqccpp" + "qq" + t + "pqp", A._fmt, layout, opA,
             opS, m, d, n, alpha, A.n_rows, A.n_cols, A.nnz, _ptr(A.vals), _ptr(A._idx0), _ptr(A._idx1), ib, 0, 0,
             D.n_rows, D.n_cols, D.family, D.major_axis, _addr(seed._c()), _addr(seed._k()), int(ro_s), int(co_s),
             beta, _ptr(B), int(ldb), st)


# ------------------------------------------------------------------------------------------------
# random sparse matrices (RandBLAS/sparse_data/random_matrix.hh) and the column partition of CSR / CSC data
def _torch_dt(np_dt):
    return {np.float32: torch.float32, np.float64: torch.float64, np.int32: torch.int32, np.int64: torch.int64}[np.dtype(np_dt).type]


def random_coo(m, n, density, state, dtype=np.float32, index_dtype=np.int64, device="cuda"):
    """RandBLAS::sparse_data::random_coo<T, sint_t>(m, n, density, state) (random_matrix.hh:290-355).
    Returns (COOMatrix on the device, next_state); the matrix equals the reference's bit for bit (rb_random_coo_*).
    `.ambiguous` on the result counts skips whose rounding could not be certified (0 in practice)."""
    _require(torch is not None, "torch available to allocate the COO arrays")
    sfx, _ = _sfx(dtype)
    ib = np.dtype(index_dtype).itemsize
    nnz, amb, nxt = ctypes.c_int64(0), ctypes.c_int64(0), (ctypes.c_uint32 * 4)()
    st = torch.cuda.current_stream().cuda_stream
    call(f"rb_random_coo_{sfx}", "qqdppqpppipppp", int(m), int(n), float(density), _addr(state._c()), _addr(state._k()), 0, 0,
         0, 0, ib, _addr(nnz), _addr(nxt), _addr(amb), st)
    k = int(nnz.value)
    vals = torch.empty(k, dtype=_torch_dt(dtype), device=device)
    rows = torch.empty(k, dtype=_torch_dt(index_dtype), device=device)
    cols = torch.empty(k, dtype=_torch_dt(index_dtype), device=device)
    if k > 0:
        call(f"rb_random_coo_{sfx}", "qqdppqpppipppp", int(m), int(n), float(density), _addr(state._c()), _addr(state._k()),
             k, _ptr(vals), _ptr(rows), _ptr(cols), ib, _addr(nnz), _addr(nxt), _addr(amb), st)
    A = COOMatrix(m, n, k, vals, rows, cols)
    A.sort = "CSR"
    A.ambiguous = int(amb.value)
    return A, RNGState(counter=list(nxt), key=state.key)


def sorted_idxs_to_compressed_ptr(n_major, idxs):
    """RandBLAS/sparse_data/base.hh:279-301 on the device."""
    ptr = _like(idxs, n_major + 1)
    nnz = int(idxs.numel()) if hasattr(idxs, "numel") else int(idxs.size)
    call("rb_sorted_idxs_to_ptr", "qqpipp", int(n_major), nnz, _ptr(idxs), np.dtype(_dtype_of(idxs)).itemsize, _ptr(ptr),
         _stream(idxs))
    return ptr


def random_csr(m, n, density, state, dtype=np.float32, index_dtype=np.int64):
    """An m x n CSR matrix with iid Bernoulli(density) pattern and N(0,1) values: the CSR form of random_coo's matrix.
    Same distribution as the reference's random_csr (random_matrix.hh:136-209) but not the same stream -- that one
    restarts its column walk per row, which is sequential by construction. Returns (CSRMatrix, next_state)."""
    coo, nxt = random_coo(m, n, density, state, dtype, index_dtype)
    A = CSRMatrix(m, n, coo.nnz, coo.vals, sorted_idxs_to_compressed_ptr(m, coo.rows), coo.cols)
    A.ambiguous = coo.ambiguous
    return A, nxt


def random_csc(m, n, density, state, dtype=np.float32, index_dtype=np.int64):
    """CSC counterpart of random_csr (random_matrix.hh:218-288): the transpose of random_coo(n, m, ...) read column-wise."""
    coo, nxt = random_coo(n, m, density, state, dtype, index_dtype)
    A = CSCMatrix(m, n, coo.nnz, coo.vals, coo.cols, sorted_idxs_to_compressed_ptr(n, coo.rows))
    A.ambiguous = coo.ambiguous
    return A, nxt


def csr_column_block(A, c0, c1):
    """Columns [c0, c1) of a device CSR matrix as a new CSRMatrix (indices shifted by -c0): the partition step in
    front of a column-sharded sketch_sparse (rb_csr_column_block)."""
    _require(A.index_base == 0, "A.index_base == IndexBase::Zero")
    vb, ib = np.dtype(_dtype_of(A.vals)).itemsize, np.dtype(_dtype_of(A.rowptr)).itemsize
    nnz = ctypes.c_int64(0)
    rowptr = _like(A.rowptr, A.n_rows + 1)
    st = _stream(A.vals, A.rowptr, A.colidxs)
    args = (A.n_rows, A.n_cols, A.nnz, _ptr(A.vals), vb, _ptr(A.rowptr), _ptr(A.colidxs), ib, int(c0), int(c1))
    call("rb_csr_column_block", "qqqpippiqqqppppp", *args, 0, 0, _ptr(rowptr), 0, _addr(nnz), st)
    k = int(nnz.value)
    vals, cols = _like(A.vals, k), _like(A.colidxs, k)
    if k > 0:
        call("rb_csr_column_block", "qqqpippiqqqppppp", *args, k, _ptr(vals), _ptr(rowptr), _ptr(cols), _addr(nnz), st)
    return CSRMatrix(A.n_rows, int(c1) - int(c0), k, vals, rowptr, cols)


def csc_column_block(A, c0, c1):
    """Columns [c0, c1) of a CSC matrix: views of vals / rowidxs plus a rebased colptr (no kernel needed)."""
    lo, hi = int(A.colptr[c0]), int(A.colptr[c1])
    return CSCMatrix(A.n_rows, int(c1) - int(c0), hi - lo, A.vals[lo:hi], A.rowidxs[lo:hi], A.colptr[c0:c1 + 1] - lo)
