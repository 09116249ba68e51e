// randblas_b200 -- header-only drop-in layer: random sparse test matrices.
// Mirrors RandBLAS/sparse_data/random_matrix.hh:136-355 of the reference (same names, template parameters and return
// types; include it explicitly, as in the reference). The entries are drawn on the device (rb_random_coo_*):
//   * random_coo reproduces the reference's sequential PhiloxStream bit for bit (rows, cols, values, returned state);
//   * random_csr / random_csc are the CSR / CSC forms of that matrix: the same distribution (iid Bernoulli(density)
//     pattern, N(0,1) values) as the reference's, but not the same stream -- theirs restarts the column walk in every
//     row, which makes a row's position in the stream depend on everything before it.
#pragma once
#include <utility>
#include "../sparse_data.hh"

namespace RandBLAS::sparse_data {

namespace detail {
inline int random_coo_c(int64_t m, int64_t n, double dens, const uint32_t* c, const uint32_t* k, int64_t cap, float* v, void* r,
                        void* cl, int ib, int64_t* nnz, uint32_t* nx, int64_t* amb) {
    return rb_random_coo_f32(m, n, dens, c, k, cap, v, r, cl, ib, nnz, nx, amb, nullptr);
}
inline int random_coo_c(int64_t m, int64_t n, double dens, const uint32_t* c, const uint32_t* k, int64_t cap, double* v, void* r,
                        void* cl, int ib, int64_t* nnz, uint32_t* nx, int64_t* amb) {
    return rb_random_coo_f64(m, n, dens, c, k, cap, v, r, cl, ib, nnz, nx, amb, nullptr);
}
}  // namespace detail

// random_matrix.hh:290-355. density in [0, 1).
template <typename T, typename sint_t = int64_t, typename RNG = RandBLAS::DefaultRNG>
std::pair<COOMatrix<T, sint_t>, RandBLAS::RNGState<RNG>> random_coo(int64_t m, int64_t n, double density,
                                                                    const RandBLAS::RNGState<RNG>& state) {
    randblas_require(density >= 0.0 && density <= 1.0);
    COOMatrix<T, sint_t> A(m, n);
    RandBLAS::RNGState<RNG> next(state);
    int64_t nnz = 0, ambiguous = 0;
    internal::check(detail::random_coo_c(m, n, density, state.counter.v, state.key.v, 0, (T*) nullptr, nullptr, nullptr,
                                         (int) sizeof(sint_t), &nnz, next.counter.v, &ambiguous),
                    __func__);
    if (nnz > 0) {
        A.reserve(nnz);
        internal::check(detail::random_coo_c(m, n, density, state.counter.v, state.key.v, nnz, A.vals, A.rows, A.cols,
                                             (int) sizeof(sint_t), &nnz, next.counter.v, &ambiguous),
                        __func__);
    }
    A.sort = NonzeroSort::CSR;
    return {std::move(A), next};
}

// random_matrix.hh:136-209
template <typename T, typename sint_t = int64_t, typename RNG = RandBLAS::DefaultRNG>
std::pair<CSRMatrix<T, sint_t>, RandBLAS::RNGState<RNG>> random_csr(int64_t m, int64_t n, double density,
                                                                    const RandBLAS::RNGState<RNG>& state) {
    auto [coo, next] = random_coo<T, sint_t, RNG>(m, n, density, state);
    CSRMatrix<T, sint_t> A(m, n);
    if (m > 0) A.rowptr = new sint_t[m + 1]{};
    if (coo.nnz > 0) {
        A.reserve(coo.nnz);
        std::copy(coo.vals, coo.vals + coo.nnz, A.vals);
        std::copy(coo.cols, coo.cols + coo.nnz, A.colidxs);
        for (int64_t e = 0; e < coo.nnz; ++e) A.rowptr[coo.rows[e] + 1] += 1;         // sorted_idxs_to_compressed_ptr
        for (int64_t i = 0; i < m; ++i) A.rowptr[i + 1] += A.rowptr[i];
    }
    return {std::move(A), next};
}

// random_matrix.hh:218-288: the transpose of random_coo(n, m, ...), read column by column
template <typename T, typename sint_t = int64_t, typename RNG = RandBLAS::DefaultRNG>
std::pair<CSCMatrix<T, sint_t>, RandBLAS::RNGState<RNG>> random_csc(int64_t m, int64_t n, double density,
                                                                    const RandBLAS::RNGState<RNG>& state) {
    auto [coo, next] = random_coo<T, sint_t, RNG>(n, m, density, state);
    CSCMatrix<T, sint_t> A(m, n);
    if (n > 0) A.colptr = new sint_t[n + 1]{};
    if (coo.nnz > 0) {
        A.reserve(coo.nnz);
        std::copy(coo.vals, coo.vals + coo.nnz, A.vals);
        std::copy(coo.cols, coo.cols + coo.nnz, A.rowidxs);
        for (int64_t e = 0; e < coo.nnz; ++e) A.colptr[coo.rows[e] + 1] += 1;
        for (int64_t j = 0; j < n; ++j) A.colptr[j + 1] += A.colptr[j];
    }
    return {std::move(A), next};
}

}  // namespace RandBLAS::sparse_data
