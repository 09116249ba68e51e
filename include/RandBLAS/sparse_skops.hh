// randblas_b200 -- header-only drop-in layer, part 3: SparseDist, SparseSkOp, fill_sparse.
// Mirrors RandBLAS/sparse_skops.hh of the reference (file:line cited per item). Short-axis-sparse operators (SASO,
// Axis::Short) are on the hot path; long-axis-sparse ones (LASO) are sampled by rb_fill_sparse_laso and applied as COO.
#pragma once
#include <algorithm>
#include <cmath>
#include <type_traits>
#include "base.hh"

namespace RandBLAS {

namespace sparse {
// sparse_skops.hh:108-114
inline double isometry_scale(Axis major_axis, int64_t vec_nnz, int64_t dim_major, int64_t dim_minor) {
    if (major_axis == Axis::Short) return std::pow((double) vec_nnz, -0.5);
    return std::sqrt(((double) dim_major) / (vec_nnz * ((double) dim_minor)));
}
}  // namespace sparse

template <typename T, typename RNG, typename sint_t>
struct SparseSkOp;

// sparse_skops.hh:131-246
struct SparseDist {
    const int64_t n_rows;
    const int64_t n_cols;
    const Axis major_axis;
    const int64_t dim_major;
    const int64_t dim_minor;
    const double isometry_scale;
    const int64_t vec_nnz;
    const int64_t full_nnz;

    SparseDist(int64_t n_rows, int64_t n_cols, int64_t vec_nnz = 4, Axis major_axis = Axis::Short)
        : n_rows(n_rows),
          n_cols(n_cols),
          major_axis(major_axis),
          dim_major((major_axis == Axis::Short) ? std::min(n_rows, n_cols) : std::max(n_rows, n_cols)),
          dim_minor(n_rows + n_cols - dim_major),
          isometry_scale(sparse::isometry_scale(major_axis, vec_nnz, dim_major, dim_minor)),
          vec_nnz(vec_nnz),
          full_nnz(vec_nnz * dim_minor) {
        randblas_require(n_rows > 0);
        randblas_require(n_cols > 0);
        randblas_require(vec_nnz > 0);
        randblas_require(vec_nnz <= dim_major);
    }

    template <typename T, typename RNG = DefaultRNG, typename sint_t = int64_t>
    SparseSkOp<T, RNG, sint_t> sample(RNGState<RNG>& seed_state) {
        return {*this, seed_state};
    }
};

// sparse_skops.hh:266-283
template <typename RNG>
inline RNGState<RNG> compute_next_state(const SparseDist& D, const RNGState<RNG>& state) {
    uint32_t next[4];
    internal::check(rb_sparse_next_state(D.n_rows, D.n_cols, D.vec_nnz, (char) D.major_axis, state.counter.v, next),
                    __func__);
    return internal::with_counter(state, next);
}

// sparse_skops.hh:289-450
template <typename T, typename RNG = DefaultRNG, typename sint_t = int64_t>
struct SparseSkOp {
    static_assert(std::is_signed<sint_t>::value && (sizeof(sint_t) == 4 || sizeof(sint_t) == 8),
                  "index type must be a signed 32- or 64-bit integer");
    using distribution_t = SparseDist;
    using state_t = RNGState<RNG>;
    using scalar_t = T;
    using index_t = sint_t;

    const SparseDist dist;
    const state_t seed_state;
    const state_t next_state;
    const int64_t n_rows;
    const int64_t n_cols;
    bool own_memory;
    int64_t nnz;          // < 0 until the operator has been sampled (sparse_skops.hh:348-360)
    T* vals;
    sint_t* rows;
    sint_t* cols;

    // standard constructor (:399-407)
    SparseSkOp(SparseDist dist, const state_t& seed_state)
        : dist(dist),
          seed_state(seed_state),
          next_state(compute_next_state(dist, seed_state)),
          n_rows(dist.n_rows),
          n_cols(dist.n_cols),
          own_memory(true),
          nnz(-1),
          vals(nullptr),
          rows(nullptr),
          cols(nullptr) {}

    // expert / view constructor (:413-428): caller-owned arrays, host or device memory
    SparseSkOp(SparseDist dist, const state_t& seed_state, const state_t& next_state, int64_t nnz, T* vals, sint_t* rows,
               sint_t* cols)
        : dist(dist),
          seed_state(seed_state),
          next_state(next_state),
          n_rows(dist.n_rows),
          n_cols(dist.n_cols),
          own_memory(false),
          nnz(nnz),
          vals(vals),
          rows(rows),
          cols(cols) {}

    SparseSkOp(SparseSkOp<T, RNG, sint_t>&& S)
        : dist(S.dist),
          seed_state(S.seed_state),
          next_state(S.next_state),
          n_rows(dist.n_rows),
          n_cols(dist.n_cols),
          own_memory(S.own_memory),
          nnz(S.nnz),
          vals(S.vals),
          rows(S.rows),
          cols(S.cols) {
        S.rows = nullptr;
        S.cols = nullptr;
        S.vals = nullptr;
        S.nnz = -1;
    }

    ~SparseSkOp() {
        if (own_memory) {
            if (rows != nullptr) delete[] rows;
            if (cols != nullptr) delete[] cols;
            if (vals != nullptr) delete[] vals;
        }
    }
};

// sparse_skops.hh:259-264: r samples of k distinct indices from {0..n-1} (sampling without replacement)
template <typename sint_t, typename RNG = DefaultRNG>
RNGState<RNG> repeated_fisher_yates(int64_t k, int64_t n, int64_t r, sint_t* samples, const RNGState<RNG>& state) {
    uint32_t next[4];
    internal::check(rb_repeated_fisher_yates(k, n, r, samples, (int) sizeof(sint_t), state.counter.v, state.key.v, next,
                                             nullptr),
                    __func__);
    return internal::with_counter(state, next);
}

// sparse_skops.hh:515-565
template <typename T, typename sint_t, typename RNG>
RNGState<RNG> fill_sparse_unpacked_nosub(const SparseDist& D, int64_t& nnz, T* vals, sint_t* rows, sint_t* cols,
                                         const RNGState<RNG>& seed_state) {
    uint32_t next[4];
    int64_t got = 0;
    if (D.major_axis == Axis::Short)     // :526-533
        internal::check(rb_fill_sparse_saso(D.n_rows, D.n_cols, D.vec_nnz, seed_state.counter.v, seed_state.key.v, vals,
                                            (int) sizeof(T), rows, cols, (int) sizeof(sint_t), &got, next, nullptr),
                        __func__);
    else                                 // :534-564
        internal::check(rb_fill_sparse_laso(D.n_rows, D.n_cols, D.vec_nnz, seed_state.counter.v, seed_state.key.v, vals,
                                            (int) sizeof(T), rows, cols, (int) sizeof(sint_t), &got, next, nullptr),
                        __func__);
    nnz = got;
    return internal::with_counter(seed_state, next);
}

// sparse_skops.hh:587-603
template <typename SparseSkOp>
void fill_sparse(SparseSkOp& S) {
    using T = typename SparseSkOp::scalar_t;
    using sint_t = typename SparseSkOp::index_t;
    const int64_t full_nnz = S.dist.full_nnz;
    if (S.own_memory) {
        if (S.rows == nullptr) S.rows = new sint_t[full_nnz];
        if (S.cols == nullptr) S.cols = new sint_t[full_nnz];
        if (S.vals == nullptr) S.vals = new T[full_nnz];
    }
    randblas_require(S.rows != nullptr);
    randblas_require(S.cols != nullptr);
    randblas_require(S.vals != nullptr);
    fill_sparse_unpacked_nosub(S.dist, S.nnz, S.vals, S.rows, S.cols, S.seed_state);
}

}  // namespace RandBLAS
