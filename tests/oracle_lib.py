"""ctypes access to the two CPU checkers (TEST INFRASTRUCTURE, never imported by randblas_b200/):

* ``port``  -- oracle/librb_oracle.so, the plain-C restatement (oracle/rb_oracle.c)
* ``ref``   -- oracle/_ref/librb_ref.so, the unmodified reference headers behind oracle/ref_capi.cc
               (present wherever it was prebuilt; it cannot be rebuilt without /root/reference)

Both are wrapped behind the same small Python surface so a test can run either.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_C = {
    "c": ctypes.c_char, "i": ctypes.c_int, "q": ctypes.c_int64, "Q": ctypes.c_uint64,
    "f": ctypes.c_float, "d": ctypes.c_double, "p": ctypes.c_void_p,
}


def _ptr(a):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "need a contiguous numpy array"
    return a.ctypes.data_as(ctypes.c_void_p)


def _call(fn, sig, args):
    assert len(sig) == len(args), (fn.__name__, len(sig), len(args))
    conv = []
    for code, a in zip(sig, args):
        if code == "p":
            conv.append(_ptr(a))
        elif code == "c":
            conv.append(ctypes.c_char(a.encode() if isinstance(a, str) else a))
        else:
            conv.append(_C[code](a))
    fn.restype = ctypes.c_int
    return fn(*conv)


def build_port():
    so = os.path.join(ORACLE_DIR, "librb_oracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("rb_oracle.c", "rb_oracle.h", "rb_oracle_typed.inc")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "librb_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def build_ref():
    """(Re)build oracle/_ref when the reference tree is present; otherwise use the prebuilt file."""
    so = os.path.join(ORACLE_DIR, "_ref", "librb_ref.so")
    if os.path.isdir("/root/reference/RandBLAS"):
        src = os.path.join(ORACLE_DIR, "ref_capi.cc")
        if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "_ref/librb_ref.so"], stdout=subprocess.DEVNULL)
    return so if os.path.exists(so) else None


def u32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint32))


def state_from_u64(k):
    """RNGState(uint64): counter 0, key = {lo32, hi32} (RandBLAS/base.hh:116-119)."""
    return u32([0, 0, 0, 0]), u32([k & 0xFFFFFFFF, (k >> 32) & 0xFFFFFFFF])


def ctr_add(ctr, n):
    v = sum(int(c) << (32 * i) for i, c in enumerate(ctr))
    v = (v + int(n)) % (1 << 128)
    return u32([(v >> (32 * i)) & 0xFFFFFFFF for i in range(4)])


class OracleError(RuntimeError):
    pass


class _Base:
    """Common numpy-level API. dtype is np.float32 / np.float64."""

    def _t(self, dtype):
        return ("f32", "f") if np.dtype(dtype) == np.float32 else ("f64", "d")

    def _check(self, rc, what):
        if rc != 0:
            raise OracleError(f"{what}: rc={rc} {self.last_error()}")


class Port(_Base):
    kind = "port"

    def __init__(self):
        self.lib = ctypes.CDLL(build_port())

    def last_error(self):
        return "argument check failed"

    def set_threads(self, n):
        self.lib.rbo_set_threads(int(n))

    def get_threads(self):
        return self.lib.rbo_get_threads()

    def philox(self, ctr, key):
        out = np.zeros(4, np.uint32)
        self.lib.rbo_philox4x32_10.restype = None
        self.lib.rbo_philox4x32_10(_ptr(u32(ctr)), _ptr(u32(key)), _ptr(out))
        return out

    def threefry(self, ctr, key):
        out = np.zeros(4, np.uint32)
        self.lib.rbo_threefry4x32_20.restype = None
        self.lib.rbo_threefry4x32_20(_ptr(u32(ctr)), _ptr(u32(key)), _ptr(out))
        return out

    def ctr_incr(self, ctr, n):
        c = u32(ctr).copy()
        self.lib.rbo_ctr_incr.restype = None
        self.lib.rbo_ctr_incr(_ptr(c), ctypes.c_uint64(n))
        return c

    def rngstate_from_u64(self, k):
        c, kk = np.zeros(4, np.uint32), np.zeros(2, np.uint32)
        self.lib.rbo_rngstate_from_u64.restype = None
        self.lib.rbo_rngstate_from_u64(ctypes.c_uint64(k), _ptr(c), _ptr(kk))
        return c, kk

    def uneg11(self, w):
        self.lib.rbo_uneg11_f32.restype = ctypes.c_float
        return np.float32(self.lib.rbo_uneg11_f32(ctypes.c_uint32(int(w))))

    def u01(self, w):
        self.lib.rbo_u01_f32.restype = ctypes.c_float
        return np.float32(self.lib.rbo_u01_f32(ctypes.c_uint32(int(w))))

    def boxmuller(self, u0, u1):
        x, y = ctypes.c_float(), ctypes.c_float()
        self.lib.rbo_boxmuller_f32.restype = None
        self.lib.rbo_boxmuller_f32(ctypes.c_uint32(int(u0)), ctypes.c_uint32(int(u1)), ctypes.byref(x), ctypes.byref(y))
        return np.float32(x.value), np.float32(y.value)

    def dense_dist_info(self, n_rows, n_cols, family="G", axis="L"):
        info = np.zeros(3, np.int64)
        iso = ctypes.c_double()
        fn = self.lib.rbo_dense_dist_info
        fn.restype = ctypes.c_int
        rc = fn(ctypes.c_int64(n_rows), ctypes.c_int64(n_cols), ctypes.c_char(family.encode()),
                ctypes.c_char(axis.encode()), _ptr(info), ctypes.byref(iso))
        self._check(rc, "dense_dist_info")
        return dict(dim_major=int(info[0]), dim_minor=int(info[1]), natural_layout=chr(int(info[2])),
                    isometry_scale=iso.value)

    def dense_next_state(self, n_rows, n_cols, family, axis, ctr, key):
        nxt = np.zeros(4, np.uint32)
        self._check(_call(self.lib.rbo_dense_next_state, "qqccpp", (n_rows, n_cols, family, axis, u32(ctr), nxt)),
                    "dense_next_state")
        return nxt

    def sparse_dist_info(self, n_rows, n_cols, vec_nnz=4, axis="S"):
        info = np.zeros(3, np.int64)
        iso = ctypes.c_double()
        fn = self.lib.rbo_sparse_dist_info
        fn.restype = ctypes.c_int
        rc = fn(ctypes.c_int64(n_rows), ctypes.c_int64(n_cols), ctypes.c_int64(vec_nnz),
                ctypes.c_char(axis.encode()), _ptr(info), ctypes.byref(iso))
        self._check(rc, "sparse_dist_info")
        return dict(dim_major=int(info[0]), dim_minor=int(info[1]), full_nnz=int(info[2]), isometry_scale=iso.value)

    def sparse_next_state(self, n_rows, n_cols, vec_nnz, axis, ctr, key):
        nxt = np.zeros(4, np.uint32)
        self._check(_call(self.lib.rbo_sparse_next_state, "qqqcpp", (n_rows, n_cols, vec_nnz, axis, u32(ctr), nxt)),
                    "sparse_next_state")
        return nxt

    def fill_dense_unpacked(self, layout, D_rows, D_cols, family, axis, n_rows, n_cols, ro_s, co_s, ctr, key, dtype):
        sfx, _ = self._t(dtype)
        buff = np.zeros(n_rows * n_cols, dtype)
        nxt = np.zeros(4, np.uint32)
        fn = getattr(self.lib, f"rbo_fill_dense_unpacked_{sfx}")
        self._check(_call(fn, "cqqccqqqqpppp", (layout, D_rows, D_cols, family, axis, n_rows, n_cols, ro_s, co_s, buff,
                                                u32(ctr), u32(key), nxt)), "fill_dense_unpacked")
        return buff, nxt

    def fill_sparse(self, D_rows, D_cols, vec_nnz, axis, ctr, key, dtype, idx_dtype=np.int64):
        nnz_full = vec_nnz * (max(D_rows, D_cols) if axis == "S" else min(D_rows, D_cols))
        vals = np.zeros(nnz_full, dtype)
        rows = np.zeros(nnz_full, idx_dtype)
        cols = np.zeros(nnz_full, idx_dtype)
        nnz = np.zeros(1, np.int64)
        nxt = np.zeros(4, np.uint32)
        rc = _call(self.lib.rbo_fill_sparse_saso if axis == "S" else self.lib.rbo_fill_sparse_laso, "qqqpppippipp",
                   (D_rows, D_cols, vec_nnz, u32(ctr), u32(key), vals, vals.itemsize, rows, cols, rows.itemsize, nnz, nxt))
        self._check(rc, "fill_sparse")
        return vals, rows, cols, int(nnz[0]), nxt

    # --- index-sampling utilities (RandBLAS/util.hh:459-560) ---
    def sample_indices_iid_uniform(self, n, k, ctr, key, idx_dtype=np.int64, rad_dtype=None):
        samples = np.full(max(k, 1), -1, idx_dtype)[:k]
        rad = None if rad_dtype is None else np.zeros(max(k, 1), rad_dtype)[:k]
        nxt = np.zeros(4, np.uint32)
        self._check(_call(self.lib.rbo_sample_indices_iid_uniform, "qqpipippp",
                          (n, k, samples, samples.itemsize, rad, 4 if rad is None else rad.itemsize, u32(ctr), u32(key),
                           nxt)), "sample_indices_iid_uniform")
        return samples, rad, nxt

    def sample_indices_iid(self, n, cdf, k, ctr, key, idx_dtype=np.int64):
        samples = np.full(max(k, 1), -1, idx_dtype)[:k]
        nxt = np.zeros(4, np.uint32)
        cdf = np.ascontiguousarray(cdf)
        self._check(_call(self.lib.rbo_sample_indices_iid, "qpiqpippp",
                          (n, cdf, cdf.itemsize, k, samples, samples.itemsize, u32(ctr), u32(key), nxt)),
                    "sample_indices_iid")
        return samples, nxt

    def weights_to_cdf(self, w, error_if_below=None):
        """Returns (cdf array, ok). ok False where the reference throws (array holds what had been written)."""
        w = np.ascontiguousarray(w).copy()
        sfx, t = self._t(w.dtype)
        if error_if_below is None:
            error_if_below = -float(np.sqrt(np.finfo(w.dtype).eps))
        rc = _call(getattr(self.lib, f"rbo_weights_to_cdf_{sfx}"), "qp" + t, (len(w), w, float(error_if_below)))
        return w, rc == 0

    def repeated_fisher_yates(self, k, n, r, ctr, key, idx_dtype=np.int64):
        samples = np.zeros(k * r, idx_dtype)
        nxt = np.zeros(4, np.uint32)
        self._check(_call(self.lib.rbo_repeated_fisher_yates, "qqqpippp",
                          (k, n, r, samples, samples.itemsize, u32(ctr), u32(key), nxt)), "repeated_fisher_yates")
        return samples, nxt

    # --- sketches. B is updated in place (numpy array) ---
    def lskge3(self, layout, opS, opA, d, n, m, alpha, dist, ctr, key, ro_s, co_s, A, lda, beta, B, ldb):
        sfx, t = self._t(B.dtype)
        fn = getattr(self.lib, f"rbo_lskge3_{sfx}")
        sig = "cccqqq" + t + "qqcc" + "pp" + "qq" + "pq" + t + "pq"
        self._check(_call(fn, sig, (layout, opS, opA, d, n, m, alpha, dist[0], dist[1], dist[2], dist[3], u32(ctr),
                                    u32(key), ro_s, co_s, A, lda, beta, B, ldb)), "lskge3")

    def rskge3(self, layout, opA, opS, m, d, n, alpha, A, lda, dist, ctr, key, ro_s, co_s, beta, B, ldb):
        sfx, t = self._t(B.dtype)
        fn = getattr(self.lib, f"rbo_rskge3_{sfx}")
        sig = "cccqqq" + t + "pq" + "qqcc" + "pp" + "qq" + t + "pq"
        self._check(_call(fn, sig, (layout, opA, opS, m, d, n, alpha, A, lda, dist[0], dist[1], dist[2], dist[3],
                                    u32(ctr), u32(key), ro_s, co_s, beta, B, ldb)), "rskge3")

    def lskges(self, layout, opS, opA, d, n, m, alpha, sdist, ctr, key, ro_s, co_s, A, lda, beta, B, ldb):
        sfx, t = self._t(B.dtype)
        assert sdist[3] == "S"
        fn = getattr(self.lib, f"rbo_lskges_{sfx}")
        sig = "cccqqq" + t + "qqq" + "pp" + "qq" + "pq" + t + "pq"
        self._check(_call(fn, sig, (layout, opS, opA, d, n, m, alpha, sdist[0], sdist[1], sdist[2], u32(ctr), u32(key),
                                    ro_s, co_s, A, lda, beta, B, ldb)), "lskges")

    def rskges(self, layout, opA, opS, m, d, n, alpha, A, lda, sdist, ctr, key, ro_s, co_s, beta, B, ldb):
        sfx, t = self._t(B.dtype)
        assert sdist[3] == "S"
        fn = getattr(self.lib, f"rbo_rskges_{sfx}")
        sig = "cccqqq" + t + "pq" + "qqq" + "pp" + "qq" + t + "pq"
        self._check(_call(fn, sig, (layout, opA, opS, m, d, n, alpha, A, lda, sdist[0], sdist[1], sdist[2], u32(ctr),
                                    u32(key), ro_s, co_s, beta, B, ldb)), "rskges")

    def lsksp3(self, fmt, layout, opS, opA, d, n, m, alpha, dist, ctr, key, ro_s, co_s, spA, beta, B, ldb,
               ro_a=0, co_a=0):
        sfx, t = self._t(B.dtype)
        fn = getattr(self.lib, f"rbo_lsksp3_{sfx}")
        A_rows, A_cols, nnz, vals, idx0, idx1 = spA
        sig = "icccqqq" + t + "qqcc" + "pp" + "qq" + "qqqppp" + "qq" + t + "pq"
        self._check(_call(fn, sig, (fmt, layout, opS, opA, d, n, m, alpha, dist[0], dist[1], dist[2], dist[3], u32(ctr),
                                    u32(key), ro_s, co_s, A_rows, A_cols, nnz, vals, idx0, idx1, ro_a, co_a, beta, B,
                                    ldb)), "lsksp3")

    def rsksp3(self, fmt, layout, opA, opS, m, d, n, alpha, spA, dist, ctr, key, ro_s, co_s, beta, B, ldb,
               ro_a=0, co_a=0):
        sfx, t = self._t(B.dtype)
        fn = getattr(self.lib, f"rbo_rsksp3_{sfx}")
        A_rows, A_cols, nnz, vals, idx0, idx1 = spA
        sig = "icccqqq" + t + "qqqppp" + "qq" + "qqcc" + "pp" + "qq" + t + "pq"
        self._check(_call(fn, sig, (fmt, layout, opA, opS, m, d, n, alpha, A_rows, A_cols, nnz, vals, idx0, idx1, ro_a,
                                    co_a, dist[0], dist[1], dist[2], dist[3], u32(ctr), u32(key), ro_s, co_s, beta, B,
                                    ldb)), "rsksp3")


class Ref(_Base):
    """The reference itself (unmodified headers) -- same Python surface as Port where it applies."""
    kind = "reference"

    def __init__(self):
        so = build_ref()
        if so is None:
            raise FileNotFoundError("oracle/_ref/librb_ref.so not built (needs /root/reference)")
        self.lib = ctypes.CDLL(so)
        self.lib.rbref_last_error.restype = ctypes.c_char_p
        self.lib.rbref_blas_config.restype = ctypes.c_char_p

    def last_error(self):
        return self.lib.rbref_last_error().decode(errors="replace")

    def blas_config(self):
        return self.lib.rbref_blas_config().decode()

    def set_threads(self, n):
        self.lib.rbref_set_threads(int(n))

    def get_threads(self):
        return self.lib.rbref_get_threads()

    def philox(self, ctr, key):
        out = np.zeros(4, np.uint32)
        self._check(_call(self.lib.rbref_philox4x32_10, "ppp", (u32(ctr), u32(key), out)), "philox")
        return out

    def ctr_incr(self, ctr, n):
        c = u32(ctr).copy()
        self._check(_call(self.lib.rbref_ctr_incr, "pQ", (c, n)), "ctr_incr")
        return c

    def rngstate_from_u64(self, k):
        c, kk = np.zeros(4, np.uint32), np.zeros(2, np.uint32)
        self._check(_call(self.lib.rbref_rngstate_from_u64, "Qpp", (k, c, kk)), "rngstate")
        return c, kk

    def uneg11_block(self, ctr, key):
        out = np.zeros(4, np.float32)
        self._check(_call(self.lib.rbref_uneg11_f32, "ppp", (u32(ctr), u32(key), out)), "uneg11")
        return out

    def boxmul_block(self, ctr, key):
        out = np.zeros(4, np.float32)
        self._check(_call(self.lib.rbref_boxmul_f32, "ppp", (u32(ctr), u32(key), out)), "boxmul")
        return out

    def dense_dist_info(self, n_rows, n_cols, family="G", axis="L"):
        info = np.zeros(3, np.int64)
        iso = np.zeros(1, np.float64)
        self._check(_call(self.lib.rbref_dense_dist_info, "qqccpp", (n_rows, n_cols, family, axis, info, iso)),
                    "dense_dist_info")
        return dict(dim_major=int(info[0]), dim_minor=int(info[1]), natural_layout=chr(int(info[2])),
                    isometry_scale=float(iso[0]))

    def dense_next_state(self, n_rows, n_cols, family, axis, ctr, key):
        nxt = np.zeros(4, np.uint32)
        self._check(_call(self.lib.rbref_dense_next_state, "qqccppp", (n_rows, n_cols, family, axis, u32(ctr), u32(key),
                                                                       nxt)), "dense_next_state")
        return nxt

    def sparse_dist_info(self, n_rows, n_cols, vec_nnz=4, axis="S"):
        info = np.zeros(3, np.int64)
        iso = np.zeros(1, np.float64)
        self._check(_call(self.lib.rbref_sparse_dist_info, "qqqcpp", (n_rows, n_cols, vec_nnz, axis, info, iso)),
                    "sparse_dist_info")
        return dict(dim_major=int(info[0]), dim_minor=int(info[1]), full_nnz=int(info[2]), isometry_scale=float(iso[0]))

    def sparse_next_state(self, n_rows, n_cols, vec_nnz, axis, ctr, key):
        nxt = np.zeros(4, np.uint32)
        self._check(_call(self.lib.rbref_sparse_next_state, "qqqcppp", (n_rows, n_cols, vec_nnz, axis, u32(ctr),
                                                                        u32(key), nxt)), "sparse_next_state")
        return nxt

    def fill_dense_unpacked(self, layout, D_rows, D_cols, family, axis, n_rows, n_cols, ro_s, co_s, ctr, key, dtype):
        sfx, _ = self._t(dtype)
        buff = np.zeros(n_rows * n_cols, dtype)
        nxt = np.zeros(4, np.uint32)
        fn = getattr(self.lib, f"rbref_fill_dense_unpacked_{sfx}")
        self._check(_call(fn, "cqqccqqqqpppp", (layout, D_rows, D_cols, family, axis, n_rows, n_cols, ro_s, co_s, buff,
                                                u32(ctr), u32(key), nxt)), "fill_dense_unpacked")
        return buff, nxt

    def fill_sparse(self, D_rows, D_cols, vec_nnz, axis, ctr, key, dtype, idx_dtype=np.int64):
        sfx, _ = self._t(dtype)
        isfx = "i32" if np.dtype(idx_dtype) == np.int32 else "i64"
        nnz_full = vec_nnz * (max(D_rows, D_cols) if axis == "S" else min(D_rows, D_cols))
        vals = np.zeros(nnz_full, dtype)
        rows = np.zeros(nnz_full, idx_dtype)
        cols = np.zeros(nnz_full, idx_dtype)
        nnz = np.zeros(1, np.int64)
        nxt = np.zeros(4, np.uint32)
        fn = getattr(self.lib, f"rbref_fill_sparse_{sfx}_{isfx}")
        self._check(_call(fn, "qqqcppppppp", (D_rows, D_cols, vec_nnz, axis, u32(ctr), u32(key), vals, rows, cols, nnz,
                                              nxt)), "fill_sparse")
        return vals, rows, cols, int(nnz[0]), nxt

    # --- index-sampling utilities (RandBLAS/util.hh:459-560), the reference's own loops ---
    def sample_indices_iid_uniform(self, n, k, ctr, key, idx_dtype=np.int64, rad_dtype=None):
        isfx = "i32" if np.dtype(idx_dtype) == np.int32 else "i64"
        sfx = "f64" if (rad_dtype is not None and np.dtype(rad_dtype) == np.float64) else "f32"
        samples = np.full(max(k, 1), -1, idx_dtype)[:k]
        rad = None if rad_dtype is None else np.zeros(max(k, 1), rad_dtype)[:k]
        nxt = np.zeros(4, np.uint32)
        fn = getattr(self.lib, f"rbref_sample_indices_iid_uniform_{sfx}_{isfx}")
        self._check(_call(fn, "qqppppp", (n, k, samples, rad, u32(ctr), u32(key), nxt)), "sample_indices_iid_uniform")
        return samples, rad, nxt

    def sample_indices_iid(self, n, cdf, k, ctr, key, idx_dtype=np.int64):
        isfx = "i32" if np.dtype(idx_dtype) == np.int32 else "i64"
        cdf = np.ascontiguousarray(cdf)
        sfx, _ = self._t(cdf.dtype)
        samples = np.full(max(k, 1), -1, idx_dtype)[:k]
        nxt = np.zeros(4, np.uint32)
        fn = getattr(self.lib, f"rbref_sample_indices_iid_{sfx}_{isfx}")
        self._check(_call(fn, "qpqpppp", (n, cdf, k, samples, u32(ctr), u32(key), nxt)), "sample_indices_iid")
        return samples, nxt

    def weights_to_cdf(self, w, error_if_below=None):
        w = np.ascontiguousarray(w).copy()
        sfx, t = self._t(w.dtype)
        if error_if_below is None:
            error_if_below = -float(np.sqrt(np.finfo(w.dtype).eps))
        rc = _call(getattr(self.lib, f"rbref_weights_to_cdf_{sfx}"), "qp" + t, (len(w), w, float(error_if_below)))
        return w, rc == 0

    def repeated_fisher_yates(self, k, n, r, ctr, key, idx_dtype=np.int64):
        isfx = "i32" if np.dtype(idx_dtype) == np.int32 else "i64"
        samples = np.zeros(k * r, idx_dtype)
        nxt = np.zeros(4, np.uint32)
        fn = getattr(self.lib, f"rbref_repeated_fisher_yates_{isfx}")
        self._check(_call(fn, "qqqpppp", (k, n, r, samples, u32(ctr), u32(key), nxt)), "repeated_fisher_yates")
        return samples, nxt

    def _skge_dense(self, left, layout, op1, op2, x, y, z, alpha, dist, ctr, key, prefill, ro_s, co_s, A, lda, beta, B,
                    ldb):
        sfx, t = self._t(B.dtype)
        fn = getattr(self.lib, f"rbref_sketch_general_dense_{sfx}")
        sig = "icccqqq" + t + "qqcc" + "pp" + "i" + "qq" + "pq" + t + "pq"
        self._check(_call(fn, sig, (left, layout, op1, op2, x, y, z, alpha, dist[0], dist[1], dist[2], dist[3],
                                    u32(ctr), u32(key), prefill, ro_s, co_s, A, lda, beta, B, ldb)), "sketch_general")

    def lskge3(self, layout, opS, opA, d, n, m, alpha, dist, ctr, key, ro_s, co_s, A, lda, beta, B, ldb, prefill=0):
        self._skge_dense(1, layout, opS, opA, d, n, m, alpha, dist, ctr, key, prefill, ro_s, co_s, A, lda, beta, B, ldb)

    def rskge3(self, layout, opA, opS, m, d, n, alpha, A, lda, dist, ctr, key, ro_s, co_s, beta, B, ldb, prefill=0):
        self._skge_dense(0, layout, opA, opS, m, d, n, alpha, dist, ctr, key, prefill, ro_s, co_s, A, lda, beta, B, ldb)

    def _skge_sparse(self, left, layout, op1, op2, x, y, z, alpha, sdist, ctr, key, prefill, ro_s, co_s, A, lda, beta,
                     B, ldb):
        sfx, t = self._t(B.dtype)
        fn = getattr(self.lib, f"rbref_sketch_general_sparse_{sfx}")
        sig = "icccqqq" + t + "qqqc" + "pp" + "i" + "qq" + "pq" + t + "pq"
        self._check(_call(fn, sig, (left, layout, op1, op2, x, y, z, alpha, sdist[0], sdist[1], sdist[2], sdist[3],
                                    u32(ctr), u32(key), prefill, ro_s, co_s, A, lda, beta, B, ldb)), "sketch_general")

    def lskges(self, layout, opS, opA, d, n, m, alpha, sdist, ctr, key, ro_s, co_s, A, lda, beta, B, ldb, prefill=0):
        self._skge_sparse(1, layout, opS, opA, d, n, m, alpha, sdist, ctr, key, prefill, ro_s, co_s, A, lda, beta, B,
                          ldb)

    def rskges(self, layout, opA, opS, m, d, n, alpha, A, lda, sdist, ctr, key, ro_s, co_s, beta, B, ldb, prefill=1):
        # prefill=1 by default: the reference's rskges with an unsampled operator throws (skge.hh:616-620)
        self._skge_sparse(0, layout, opA, opS, m, d, n, alpha, sdist, ctr, key, prefill, ro_s, co_s, A, lda, beta, B,
                          ldb)

    def _sksp(self, left, fmt, layout, op1, op2, x, y, z, alpha, dist, ctr, key, prefill, ro_s, co_s, spA, beta, B, ldb):
        sfx, t = self._t(B.dtype)
        fn = getattr(self.lib, f"rbref_sketch_sparse_{sfx}")
        A_rows, A_cols, nnz, vals, idx0, idx1 = spA
        sig = "iicccqqq" + t + "qqcc" + "pp" + "i" + "qq" + "qqqppp" + t + "pq"
        self._check(_call(fn, sig, (left, fmt, layout, op1, op2, x, y, z, alpha, dist[0], dist[1], dist[2], dist[3],
                                    u32(ctr), u32(key), prefill, ro_s, co_s, A_rows, A_cols, nnz, vals, idx0, idx1,
                                    beta, B, ldb)), "sketch_sparse")

    def lsksp3(self, fmt, layout, opS, opA, d, n, m, alpha, dist, ctr, key, ro_s, co_s, spA, beta, B, ldb, prefill=0):
        self._sksp(1, fmt, layout, opS, opA, d, n, m, alpha, dist, ctr, key, prefill, ro_s, co_s, spA, beta, B, ldb)

    def rsksp3(self, fmt, layout, opA, opS, m, d, n, alpha, spA, dist, ctr, key, ro_s, co_s, beta, B, ldb, prefill=0):
        self._sksp(0, fmt, layout, opA, opS, m, d, n, alpha, dist, ctr, key, prefill, ro_s, co_s, spA, beta, B, ldb)

    # --- sparse data: public spmm, conversions, random matrices (the reference's own code; int64 indices) ---
    def spmm(self, side_left, fmt, layout, op1, op2, x, y, z, alpha, spA, ro_a, co_a, B, ldb, beta, C, ldc):
        """left_spmm(layout, opA, opB, d, n, m, ...) for side_left = 1; right_spmm(layout, opA(dense), opB(sparse),
        m, d, n, ...) for side_left = 0 (spmm_dispatch.hh:52-219). C is updated in place."""
        sfx, t = self._t(C.dtype)
        A_rows, A_cols, nnz, vals, idx0, idx1 = spA
        fn = getattr(self.lib, f"rbref_spmm_{sfx}")
        sig = "iicccqqq" + t + "qqqppp" + "qq" + "pq" + t + "pq"
        self._check(_call(fn, sig, (side_left, fmt, layout, op1, op2, x, y, z, alpha, A_rows, A_cols, nnz, vals, idx0, idx1,
                                    ro_a, co_a, B, ldb, beta, C, ldc)), "spmm")

    def coo_to_compressed(self, to_csc, n_rows, n_cols, vals, rows, cols):
        nnz = len(vals)
        sfx, _ = self._t(vals.dtype)
        ov, oi = np.zeros(nnz, vals.dtype), np.zeros(nnz, np.int64)
        op = np.zeros((n_cols if to_csc else n_rows) + 1, np.int64)
        fn = getattr(self.lib, f"rbref_coo_to_compressed_{sfx}")
        self._check(_call(fn, "iqqqpppppp", (to_csc, n_rows, n_cols, nnz, np.ascontiguousarray(vals),
                                             np.ascontiguousarray(rows, np.int64), np.ascontiguousarray(cols, np.int64),
                                             ov, oi, op)), "coo_to_compressed")
        return ov, oi, op

    def compressed_to_coo(self, from_csc, n_rows, n_cols, vals, idx, ptr):
        nnz = len(vals)
        sfx, _ = self._t(vals.dtype)
        out = np.zeros(nnz, np.int64)
        fn = getattr(self.lib, f"rbref_compressed_to_coo_{sfx}")
        self._check(_call(fn, "iqqqpppp", (from_csc, n_rows, n_cols, nnz, np.ascontiguousarray(vals),
                                           np.ascontiguousarray(idx, np.int64), np.ascontiguousarray(ptr, np.int64), out)),
                    "compressed_to_coo")
        return out

    def random_sparse(self, which, m, n, density, ctr, key, dtype=np.float32):
        """random_csr (which=0) / random_csc (1) / random_coo (2), random_matrix.hh:136-355.
        Returns (vals, idx0, idx1, nnz, next_ctr) in the (vals, idx0, idx1) convention of the C ABI."""
        sfx, _ = self._t(dtype)
        fn = getattr(self.lib, f"rbref_random_sparse_{sfx}")
        nnz, nxt = np.zeros(1, np.int64), np.zeros(4, np.uint32)
        n0 = lambda k: (m + 1) if which == 0 else k
        n1 = lambda k: (n + 1) if which == 1 else k
        v, i0, i1 = np.zeros(1, dtype), np.zeros(n0(1), np.int64), np.zeros(n1(1), np.int64)
        self._check(_call(fn, "iqqdppqppppp", (which, m, n, float(density), u32(ctr), u32(key), 0, v, i0, i1, nnz, nxt)),
                    "random_sparse")
        k = int(nnz[0])
        v, i0, i1 = np.zeros(max(k, 1), dtype), np.zeros(n0(max(k, 1)), np.int64), np.zeros(n1(max(k, 1)), np.int64)
        self._check(_call(fn, "iqqdppqppppp", (which, m, n, float(density), u32(ctr), u32(key), max(k, 1), v, i0, i1, nnz,
                                               nxt)), "random_sparse")
        return v[:k], (i0 if which == 0 else i0[:k]), (i1 if which == 1 else i1[:k]), k, nxt

    def sketch_vector_dense(self, opS, d, m, alpha, dist, ctr, key, ro_s, co_s, x, incx, beta, y, incy):
        sfx, t = self._t(y.dtype)
        fn = getattr(self.lib, f"rbref_sketch_vector_dense_{sfx}")
        sig = "cqq" + t + "qqcc" + "pp" + "qq" + "pq" + t + "pq"
        self._check(_call(fn, sig, (opS, d, m, alpha, dist[0], dist[1], dist[2], dist[3], u32(ctr), u32(key), ro_s,
                                    co_s, x, incx, beta, y, incy)), "sketch_vector")

    def sketch_vector_sparse(self, opS, d, m, alpha, sdist, ctr, key, ro_s, co_s, x, incx, beta, y, incy):
        sfx, t = self._t(y.dtype)
        fn = getattr(self.lib, f"rbref_sketch_vector_sparse_{sfx}")
        sig = "cqq" + t + "qqqc" + "pp" + "qq" + "pq" + t + "pq"
        self._check(_call(fn, sig, (opS, d, m, alpha, sdist[0], sdist[1], sdist[2], sdist[3], u32(ctr), u32(key), ro_s,
                                    co_s, x, incx, beta, y, incy)), "sketch_vector")


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        _port = Port()
    return _port


def ref():
    """The compiled reference, or None when oracle/_ref was never built for this checkout."""
    global _ref
    if _ref is None:
        try:
            _ref = Ref()
        except (FileNotFoundError, OSError):
            return None
    return _ref
