// randblas_b200 -- header-only drop-in layer, part 5: the GEMM-like sketching drivers.
// Mirrors RandBLAS/skge.hh (sketch_general, 8 overloads), skve.hh (sketch_vector) and sparse_data/sksp.hh
// (sketch_sparse) of the reference. Each call forwards to exactly one C-ABI entry point; A, B, x, y and the
// operator arrays may be host memory (staged inside the call, results are back when the call returns) or
// device memory (used in place; the call is ordered on the default stream and synchronised before returning
// to keep the reference's synchronous semantics).
#pragma once
#include "dense_skops.hh"
#include "sparse_data.hh"
#include "sparse_skops.hh"

namespace RandBLAS {

namespace internal {
#define RB_DEF_T(T, sfx)                                                                                                 \
    template <typename RNG>                                                                                              \
    inline int lskge3_c(char l, char oS, char oA, int64_t d, int64_t n, int64_t m, T alpha, const DenseSkOp<T, RNG>& S,  \
                        int64_t ro_s, int64_t co_s, const T* A, int64_t lda, T beta, T* B, int64_t ldb) {                \
        return rb_lskge3_##sfx(l, oS, oA, d, n, m, alpha, S.dist.n_rows, S.dist.n_cols, (char) S.dist.family,           \
                               (char) S.dist.major_axis, S.seed_state.counter.v, S.seed_state.key.v, S.buff, ro_s, co_s, \
                               A, lda, beta, B, ldb, nullptr);                                                           \
    }                                                                                                                    \
    template <typename RNG>                                                                                              \
    inline int rskge3_c(char l, char oA, char oS, int64_t m, int64_t d, int64_t n, T alpha, const T* A, int64_t lda,     \
                        const DenseSkOp<T, RNG>& S, int64_t ro_s, int64_t co_s, T beta, T* B, int64_t ldb) {             \
        return rb_rskge3_##sfx(l, oA, oS, m, d, n, alpha, A, lda, S.dist.n_rows, S.dist.n_cols, (char) S.dist.family,    \
                               (char) S.dist.major_axis, S.seed_state.counter.v, S.seed_state.key.v, S.buff, ro_s, co_s, \
                               beta, B, ldb, nullptr);                                                                   \
    }                                                                                                                    \
    template <typename RNG, typename I>                                                                                  \
    inline int skges_c(bool left, char l, char oS, char oA, int64_t d, int64_t n, int64_t m, T alpha,                    \
                       const SparseSkOp<T, RNG, I>& S, int64_t ro_s, int64_t co_s, const T* A, int64_t lda, T beta, T* B, \
                       int64_t ldb) {                                                                                    \
        if (S.nnz >= 0) /* already sampled: the COO view of the operator (skge.hh:489-490) */                            \
            return rb_coo_apply_##sfx(left ? 1 : 0, l, oS, oA, d, n, m, alpha, S.n_rows, S.n_cols, S.nnz, S.vals, S.rows, \
                                      S.cols, (int) sizeof(I), ro_s, co_s, A, lda, beta, B, ldb, nullptr);               \
        if (left)                                                                                                        \
            return rb_lskges_##sfx(l, oS, oA, d, n, m, alpha, S.dist.n_rows, S.dist.n_cols, S.dist.vec_nnz,              \
                                   S.seed_state.counter.v, S.seed_state.key.v, ro_s, co_s, A, lda, beta, B, ldb, nullptr); \
        return rb_rskges_##sfx(l, oA, oS, m, d, n, alpha, A, lda, S.dist.n_rows, S.dist.n_cols, S.dist.vec_nnz,          \
                               S.seed_state.counter.v, S.seed_state.key.v, ro_s, co_s, beta, B, ldb, nullptr);           \
    }                                                                                                                    \
    template <typename RNG, typename SpMat>                                                                              \
    inline int sksp3_c(bool left, char l, char oS, char oA, int64_t d, int64_t n, int64_t m, T alpha,                    \
                       const DenseSkOp<T, RNG>& S, int64_t ro_s, int64_t co_s, const SpMat& A, int64_t ro_a, int64_t co_a, \
                       T beta, T* B, int64_t ldb) {                                                                      \
        int fmt; const void* i0; const void* i1;                                                                         \
        sp_arrays(A, fmt, i0, i1);                                                                                       \
        const int ib = (int) sizeof(typename SpMat::index_t);                                                            \
        if (left)                                                                                                        \
            return rb_lsksp3_##sfx(fmt, l, oS, oA, d, n, m, alpha, S.dist.n_rows, S.dist.n_cols, (char) S.dist.family,   \
                                   (char) S.dist.major_axis, S.seed_state.counter.v, S.seed_state.key.v, ro_s, co_s,     \
                                   A.n_rows, A.n_cols, A.nnz, A.vals, i0, i1, ib, ro_a, co_a, beta, B, ldb, nullptr);    \
        return rb_rsksp3_##sfx(fmt, l, oA, oS, m, d, n, alpha, A.n_rows, A.n_cols, A.nnz, A.vals, i0, i1, ib, ro_a, co_a, \
                               S.dist.n_rows, S.dist.n_cols, (char) S.dist.family, (char) S.dist.major_axis,            \
                               S.seed_state.counter.v, S.seed_state.key.v, ro_s, co_s, beta, B, ldb, nullptr);           \
    }
RB_DEF_T(float, f32)
RB_DEF_T(double, f64)
#undef RB_DEF_T

inline int spmm_c(int left, int fmt, char l, char oA, char oB, int64_t d, int64_t n, int64_t m, float alpha, int64_t ar,
                  int64_t ac, int64_t nnz, const float* vals, const void* i0, const void* i1, int ib, int64_t ro,
                  int64_t co, const float* B, int64_t ldb, float beta, float* C, int64_t ldc) {
    return rb_spmm_f32(left, fmt, l, oA, oB, d, n, m, alpha, ar, ac, nnz, vals, i0, i1, ib, ro, co, B, ldb, beta, C, ldc,
                       nullptr);
}
inline int spmm_c(int left, int fmt, char l, char oA, char oB, int64_t d, int64_t n, int64_t m, double alpha, int64_t ar,
                  int64_t ac, int64_t nnz, const double* vals, const void* i0, const void* i1, int ib, int64_t ro,
                  int64_t co, const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
    return rb_spmm_f64(left, fmt, l, oA, oB, d, n, m, alpha, ar, ac, nnz, vals, i0, i1, ib, ro, co, B, ldb, beta, C, ldc,
                       nullptr);
}

// device-pointer calls are stream-ordered; keep the reference's synchronous semantics
inline void finish(int rc, const char* func) {
    check(rc, func);
    check(rb_sync_stream(nullptr), func);
}
}  // namespace internal

// ===================================================================== sketch_general, operator on the left
// skge.hh:799-821: B = alpha * op(S[ro_s:, co_s:]) * op(A) + beta * B, dense operator
template <typename T, typename RNG>
inline void sketch_general(blas::Layout layout, blas::Op opS, blas::Op opA, int64_t d, int64_t n, int64_t m, T alpha,
                           const DenseSkOp<T, RNG>& S, int64_t ro_s, int64_t co_s, const T* A, int64_t lda, T beta, T* B,
                           int64_t ldb) {
    internal::finish(internal::lskge3_c(internal::to_char(layout), internal::to_char(opS), internal::to_char(opA), d, n, m,
                                        alpha, S, ro_s, co_s, A, lda, beta, B, ldb),
                     __func__);
}
// skge.hh:775-797: sparse operator
template <typename T, typename RNG, typename sint_t>
inline void sketch_general(blas::Layout layout, blas::Op opS, blas::Op opA, int64_t d, int64_t n, int64_t m, T alpha,
                           const SparseSkOp<T, RNG, sint_t>& S, int64_t ro_s, int64_t co_s, const T* A, int64_t lda, T beta,
                           T* B, int64_t ldb) {
    if (S.dist.major_axis != Axis::Short && S.nnz < 0) {
        // LASO, not yet sampled: a temporary is sampled and dropped, as in the reference (skge.hh:483-488)
        SparseSkOp<T, RNG, sint_t> tmp(S.dist, S.seed_state);
        fill_sparse(tmp);
        return sketch_general(layout, opS, opA, d, n, m, alpha, tmp, ro_s, co_s, A, lda, beta, B, ldb);
    }
    internal::finish(internal::skges_c(true, internal::to_char(layout), internal::to_char(opS), internal::to_char(opA), d, n,
                                       m, alpha, S, ro_s, co_s, A, lda, beta, B, ldb),
                     __func__);
}
// skge.hh:1073-1097: full operator, dimensions checked
template <typename T, typename SKOP>
inline void sketch_general(blas::Layout layout, blas::Op opS, blas::Op opA, int64_t d, int64_t n, int64_t m, T alpha,
                           const SKOP& S, const T* A, int64_t lda, T beta, T* B, int64_t ldb) {
    if (opS == blas::Op::NoTrans) {
        randblas_require(S.n_rows == d);
        randblas_require(S.n_cols == m);
    } else {
        randblas_require(S.n_rows == m);
        randblas_require(S.n_cols == d);
    }
    return sketch_general(layout, opS, opA, d, n, m, alpha, S, (int64_t) 0, (int64_t) 0, A, lda, beta, B, ldb);
}

// ==================================================================== sketch_general, operator on the right
// skge.hh:947-968: B = alpha * op(A) * op(S[ro_s:, co_s:]) + beta * B, dense operator
template <typename T, typename RNG>
inline void sketch_general(blas::Layout layout, blas::Op opA, blas::Op opS, int64_t m, int64_t d, int64_t n, T alpha,
                           const T* A, int64_t lda, const DenseSkOp<T, RNG>& S, int64_t ro_s, int64_t co_s, T beta, T* B,
                           int64_t ldb) {
    internal::finish(internal::rskge3_c(internal::to_char(layout), internal::to_char(opA), internal::to_char(opS), m, d, n,
                                        alpha, A, lda, S, ro_s, co_s, beta, B, ldb),
                     __func__);
}
// skge.hh:971-992: sparse operator. (The reference's rskges falls through after sampling an unsampled operator,
// skge.hh:616-620, and throws; here an unsampled operator is applied like on the left.)
template <typename T, typename RNG, typename sint_t>
inline void sketch_general(blas::Layout layout, blas::Op opA, blas::Op opS, int64_t m, int64_t d, int64_t n, T alpha,
                           const T* A, int64_t lda, const SparseSkOp<T, RNG, sint_t>& S, int64_t ro_s, int64_t co_s, T beta,
                           T* B, int64_t ldb) {
    if (S.dist.major_axis != Axis::Short && S.nnz < 0) {
        SparseSkOp<T, RNG, sint_t> tmp(S.dist, S.seed_state);
        fill_sparse(tmp);
        return sketch_general(layout, opA, opS, m, d, n, alpha, A, lda, tmp, ro_s, co_s, beta, B, ldb);
    }
    internal::finish(internal::skges_c(false, internal::to_char(layout), internal::to_char(opS), internal::to_char(opA), d, n,
                                       m, alpha, S, ro_s, co_s, A, lda, beta, B, ldb),
                     __func__);
}
// skge.hh:1175-1199
template <typename T, typename SKOP>
inline void sketch_general(blas::Layout layout, blas::Op opA, blas::Op opS, int64_t m, int64_t d, int64_t n, T alpha,
                           const T* A, int64_t lda, const SKOP& S, T beta, T* B, int64_t ldb) {
    if (opS == blas::Op::NoTrans) {
        randblas_require(S.n_rows == n);
        randblas_require(S.n_cols == d);
    } else {
        randblas_require(S.n_rows == d);
        randblas_require(S.n_cols == n);
    }
    return sketch_general(layout, opA, opS, m, d, n, alpha, A, lda, S, (int64_t) 0, (int64_t) 0, beta, B, ldb);
}

// ============================================================================================ sketch_vector
// skve.hh:141-164: y = alpha * op(S[ro_s:, co_s:]) * x + beta * y
template <typename T, typename SKOP>
inline void sketch_vector(blas::Op opS, int64_t d, int64_t m, T alpha, const SKOP& S, int64_t ro_s, int64_t co_s,
                          const T* x, int64_t incx, T beta, T* y, int64_t incy) {
    int64_t _d = d, _m = m;
    if (opS == blas::Op::Trans) { _d = m; _m = d; }
    return sketch_general(blas::Layout::RowMajor, opS, blas::Op::NoTrans, _d, (int64_t) 1, _m, alpha, S, ro_s, co_s, x, incx,
                          beta, y, incy);
}
// skve.hh:233-246
template <typename T, typename SKOP>
inline void sketch_vector(blas::Op opS, T alpha, const SKOP& S, const T* x, int64_t incx, T beta, T* y, int64_t incy) {
    return sketch_vector(opS, S.n_rows, S.n_cols, alpha, S, (int64_t) 0, (int64_t) 0, x, incx, beta, y, incy);
}

// ============================================================================================ sketch_sparse
// sksp.hh:418-437: B = alpha * op(S[ro_s:, co_s:]) * op(A_sparse) + beta * B
template <typename SpMat, typename DenseSkOp, typename T = typename DenseSkOp::scalar_t>
inline void sketch_sparse(blas::Layout layout, blas::Op opS, blas::Op opA, int64_t d, int64_t n, int64_t m, T alpha,
                          const DenseSkOp& S, int64_t ro_s, int64_t co_s, const SpMat& A, T beta, T* B, int64_t ldb) {
    randblas_require(A.index_base == IndexBase::Zero);     // spmm_dispatch.hh:92
    internal::finish(internal::sksp3_c(true, internal::to_char(layout), internal::to_char(opS), internal::to_char(opA), d, n,
                                       m, alpha, S, ro_s, co_s, A, (int64_t) 0, (int64_t) 0, beta, B, ldb),
                     __func__);
}
// sksp.hh:520-539: B = alpha * op(A_sparse) * op(S[ro_s:, co_s:]) + beta * B
template <typename SpMat, typename DenseSkOp, typename T = typename DenseSkOp::scalar_t>
inline void sketch_sparse(blas::Layout layout, blas::Op opA, blas::Op opS, int64_t m, int64_t d, int64_t n, T alpha,
                          const SpMat& A, const DenseSkOp& S, int64_t ro_s, int64_t co_s, T beta, T* B, int64_t ldb) {
    randblas_require(A.index_base == IndexBase::Zero);
    internal::finish(internal::sksp3_c(false, internal::to_char(layout), internal::to_char(opS), internal::to_char(opA), d, n,
                                       m, alpha, S, ro_s, co_s, A, (int64_t) 0, (int64_t) 0, beta, B, ldb),
                     __func__);
}

// ============================================================================================ sketch_symmetric
namespace util {
// util.hh:128-148
inline void require_symmetric(blas::Layout layout, const float* A, int64_t n, int64_t lda, float tol) {
    internal::check(rb_require_symmetric_f32(internal::to_char(layout), A, n, lda, tol, nullptr), __func__);
}
inline void require_symmetric(blas::Layout layout, const double* A, int64_t n, int64_t lda, double tol) {
    internal::check(rb_require_symmetric_f64(internal::to_char(layout), A, n, lda, tol, nullptr), __func__);
}
}  // namespace util
// sksy.hh:159-176: B(n x d) = alpha * A * S[ro_s:, co_s:] + beta * B, A symmetric, stored as a general matrix
template <typename SKOP, typename T = typename SKOP::scalar_t>
inline void sketch_symmetric(blas::Layout layout, int64_t n, int64_t d, T alpha, const T* A, int64_t lda, const SKOP& S,
                             int64_t ro_s, int64_t co_s, T beta, T* B, int64_t ldb, T sym_check_tol = 0) {
    util::require_symmetric(layout, A, n, lda, sym_check_tol);
    sketch_general(layout, blas::Op::NoTrans, blas::Op::NoTrans, n, d, n, alpha, A, lda, S, ro_s, co_s, beta, B, ldb);
}
// sksy.hh:294-312: B(d x n) = alpha * S[ro_s:, co_s:] * A + beta * B
template <typename SKOP, typename T = typename SKOP::scalar_t>
inline void sketch_symmetric(blas::Layout layout, int64_t d, int64_t n, T alpha, const SKOP& S, int64_t ro_s, int64_t co_s,
                             const T* A, int64_t lda, T beta, T* B, int64_t ldb, T sym_check_tol = 0) {
    util::require_symmetric(layout, A, n, lda, sym_check_tol);
    sketch_general(layout, blas::Op::NoTrans, blas::Op::NoTrans, d, n, n, alpha, S, ro_s, co_s, A, lda, beta, B, ldb);
}

// ======================================================================= sparse data times dense matrix
namespace sparse_data {
// spmm_dispatch.hh:52-178: C(d x n) = alpha * op(A_sp[ro_a:, co_a:]) * op(B) + beta * C
template <typename SpMat, typename T = typename SpMat::scalar_t>
inline void left_spmm(blas::Layout layout, blas::Op opA, blas::Op opB, int64_t d, int64_t n, int64_t m, T alpha,
                      const SpMat& A, int64_t ro_a, int64_t co_a, const T* B, int64_t ldb, T beta, T* C, int64_t ldc) {
    randblas_require(A.index_base == IndexBase::Zero);     // :92
    int fmt; const void* i0; const void* i1;
    internal::sp_arrays(A, fmt, i0, i1);
    internal::finish(internal::spmm_c(1, fmt, internal::to_char(layout), internal::to_char(opA), internal::to_char(opB), d, n,
                                      m, alpha, A.n_rows, A.n_cols, A.nnz, A.vals, i0, i1,
                                      (int) sizeof(typename SpMat::index_t), ro_a, co_a, B, ldb, beta, C, ldc),
                     __func__);
}
// spmm_dispatch.hh:180-219: C(m x d) = alpha * op(A) * op(B_sp[i_off:, j_off:]) + beta * C
template <typename SpMat, typename T = typename SpMat::scalar_t>
inline void right_spmm(blas::Layout layout, blas::Op opA, blas::Op opB, int64_t m, int64_t d, int64_t n, T alpha,
                       const T* A, int64_t lda, const SpMat& B, int64_t i_off, int64_t j_off, T beta, T* C, int64_t ldc) {
    randblas_require(B.index_base == IndexBase::Zero);
    int fmt; const void* i0; const void* i1;
    internal::sp_arrays(B, fmt, i0, i1);
    internal::finish(internal::spmm_c(0, fmt, internal::to_char(layout), internal::to_char(opB), internal::to_char(opA), d, n,
                                      m, alpha, B.n_rows, B.n_cols, B.nnz, B.vals, i0, i1,
                                      (int) sizeof(typename SpMat::index_t), i_off, j_off, A, lda, beta, C, ldc),
                     __func__);
}
}  // namespace sparse_data

}  // namespace RandBLAS
