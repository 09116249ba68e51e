// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Minimal BLAS++-compatible surface (enums + gemm/scal/copy/axpy/gemv/nrm2/dot) over the
// CBLAS symbols of scipy's bundled OpenBLAS (`scipy_cblas_*`, LP64). BLAS++
// (github.com/icl-utk-edu/blaspp, unpinned HEAD in the reference CI) is an un-vendored
// dependency of the reference; call sites: RandBLAS/skge.hh:200,353, dense_skops.hh:589,599,
// util.hh:60,470, sparse_data/csc_spmm_impl.hh:201, csr_spmm_impl.hh:150.
#pragma once
#ifndef BLAS_HH
#define BLAS_HH
#include <cstdint>
#include <cmath>
#include <complex>
#include <limits>
#include <sstream>
#include <vector>
#include <numeric>
#include <algorithm>
#include <stdexcept>
#include <string>

extern "C" {
void scipy_cblas_sgemm(int layout, int ta, int tb, int m, int n, int k, float alpha, const float* A, int lda,
                       const float* B, int ldb, float beta, float* C, int ldc);
void scipy_cblas_dgemm(int layout, int ta, int tb, int m, int n, int k, double alpha, const double* A, int lda,
                       const double* B, int ldb, double beta, double* C, int ldc);
void scipy_cblas_sgemv(int layout, int ta, int m, int n, float alpha, const float* A, int lda, const float* x,
                       int incx, float beta, float* y, int incy);
void scipy_cblas_dgemv(int layout, int ta, int m, int n, double alpha, const double* A, int lda, const double* x,
                       int incx, double beta, double* y, int incy);
void scipy_cblas_sscal(int n, float a, float* x, int incx);
void scipy_cblas_dscal(int n, double a, double* x, int incx);
void scipy_cblas_scopy(int n, const float* x, int incx, float* y, int incy);
void scipy_cblas_dcopy(int n, const double* x, int incx, double* y, int incy);
void scipy_cblas_saxpy(int n, float a, const float* x, int incx, float* y, int incy);
void scipy_cblas_daxpy(int n, double a, const double* x, int incx, double* y, int incy);
float scipy_cblas_snrm2(int n, const float* x, int incx);
double scipy_cblas_dnrm2(int n, const double* x, int incx);
float scipy_cblas_sdot(int n, const float* x, int incx, const float* y, int incy);
double scipy_cblas_ddot(int n, const double* x, int incx, const double* y, int incy);
void scipy_openblas_set_num_threads(int n);
int scipy_openblas_get_num_threads(void);
char* scipy_openblas_get_config(void);
}

namespace blas {

enum class Layout : char { ColMajor = 'C', RowMajor = 'R' };
enum class Op     : char { NoTrans = 'N', Trans = 'T', ConjTrans = 'C' };
enum class Uplo   : char { Upper = 'U', Lower = 'L', General = 'G' };
enum class Diag   : char { NonUnit = 'N', Unit = 'U' };
enum class Side   : char { Left = 'L', Right = 'R' };

namespace detail {
inline int cl(Layout l) { return l == Layout::RowMajor ? 101 : 102; }
inline int co(Op o) { return o == Op::NoTrans ? 111 : (o == Op::Trans ? 112 : 113); }
inline int ci(int64_t v) {
    if (v > std::numeric_limits<int>::max() || v < std::numeric_limits<int>::min())
        throw std::overflow_error("blas shim: dimension exceeds LP64 int");
    return (int) v;
}
// scal/copy/axpy on vectors longer than INT_MAX are chunked (unit or any stride)
template <typename F> inline void chunked(int64_t n, F f) {
    const int64_t step = 1ll << 30;
    for (int64_t s = 0; s < n; s += step) f(s, (int) std::min<int64_t>(step, n - s));
}
}

inline void gemm(Layout l, Op ta, Op tb, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda,
                 const float* B, int64_t ldb, float beta, float* C, int64_t ldc) {
    using namespace detail;
    scipy_cblas_sgemm(cl(l), co(ta), co(tb), ci(m), ci(n), ci(k), alpha, A, ci(lda), B, ci(ldb), beta, C, ci(ldc));
}
inline void gemm(Layout l, Op ta, Op tb, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                 const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
    using namespace detail;
    scipy_cblas_dgemm(cl(l), co(ta), co(tb), ci(m), ci(n), ci(k), alpha, A, ci(lda), B, ci(ldb), beta, C, ci(ldc));
}
inline void gemv(Layout l, Op ta, int64_t m, int64_t n, float alpha, const float* A, int64_t lda, const float* x,
                 int64_t incx, float beta, float* y, int64_t incy) {
    using namespace detail;
    scipy_cblas_sgemv(cl(l), co(ta), ci(m), ci(n), alpha, A, ci(lda), x, ci(incx), beta, y, ci(incy));
}
inline void gemv(Layout l, Op ta, int64_t m, int64_t n, double alpha, const double* A, int64_t lda, const double* x,
                 int64_t incx, double beta, double* y, int64_t incy) {
    using namespace detail;
    scipy_cblas_dgemv(cl(l), co(ta), ci(m), ci(n), alpha, A, ci(lda), x, ci(incx), beta, y, ci(incy));
}
inline void scal(int64_t n, float a, float* x, int64_t incx) {
    detail::chunked(n, [&](int64_t s, int c) { scipy_cblas_sscal(c, a, x + s * incx, detail::ci(incx)); });
}
inline void scal(int64_t n, double a, double* x, int64_t incx) {
    detail::chunked(n, [&](int64_t s, int c) { scipy_cblas_dscal(c, a, x + s * incx, detail::ci(incx)); });
}
inline void copy(int64_t n, const float* x, int64_t incx, float* y, int64_t incy) {
    detail::chunked(n, [&](int64_t s, int c) { scipy_cblas_scopy(c, x + s * incx, detail::ci(incx), y + s * incy, detail::ci(incy)); });
}
inline void copy(int64_t n, const double* x, int64_t incx, double* y, int64_t incy) {
    detail::chunked(n, [&](int64_t s, int c) { scipy_cblas_dcopy(c, x + s * incx, detail::ci(incx), y + s * incy, detail::ci(incy)); });
}
inline void axpy(int64_t n, float a, const float* x, int64_t incx, float* y, int64_t incy) {
    detail::chunked(n, [&](int64_t s, int c) { scipy_cblas_saxpy(c, a, x + s * incx, detail::ci(incx), y + s * incy, detail::ci(incy)); });
}
inline void axpy(int64_t n, double a, const double* x, int64_t incx, double* y, int64_t incy) {
    detail::chunked(n, [&](int64_t s, int c) { scipy_cblas_daxpy(c, a, x + s * incx, detail::ci(incx), y + s * incy, detail::ci(incy)); });
}
inline float nrm2(int64_t n, const float* x, int64_t incx) { return scipy_cblas_snrm2(detail::ci(n), x, detail::ci(incx)); }
inline double nrm2(int64_t n, const double* x, int64_t incx) { return scipy_cblas_dnrm2(detail::ci(n), x, detail::ci(incx)); }
inline float dot(int64_t n, const float* x, int64_t incx, const float* y, int64_t incy) {
    return scipy_cblas_sdot(detail::ci(n), x, detail::ci(incx), y, detail::ci(incy));
}
inline double dot(int64_t n, const double* x, int64_t incx, const double* y, int64_t incy) {
    return scipy_cblas_ddot(detail::ci(n), x, detail::ci(incx), y, detail::ci(incy));
}
} // namespace blas
#endif // BLAS_HH
