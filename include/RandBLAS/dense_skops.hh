// randblas_b200 -- header-only drop-in layer, part 2: DenseDist, DenseSkOp, fill_dense.
// Mirrors RandBLAS/dense_skops.hh of the reference (file:line cited per item).
#pragma once
#include <algorithm>
#include <cmath>
#include "base.hh"

namespace RandBLAS {

// dense_skops.hh:215-224
enum class ScalarDist : char { Gaussian = 'G', Uniform = 'U' };

namespace dense {
// dense_skops.hh:187-199
inline blas::Layout natural_layout(Axis major_axis, int64_t n_rows, int64_t n_cols) {
    const bool is_wide = n_rows < n_cols, fa_long = major_axis == Axis::Long;
    if (is_wide && fa_long) return blas::Layout::RowMajor;
    if (is_wide) return blas::Layout::ColMajor;
    if (fa_long) return blas::Layout::ColMajor;
    return blas::Layout::RowMajor;
}
}  // namespace dense

template <typename T, typename RNG>
struct DenseSkOp;

// dense_skops.hh:231-350
struct DenseDist {
    const int64_t n_rows;
    const int64_t n_cols;
    const Axis major_axis;
    const int64_t dim_major;
    const int64_t dim_minor;
    const double isometry_scale;
    const ScalarDist family;
    const blas::Layout natural_layout;

    DenseDist(int64_t n_rows, int64_t n_cols, ScalarDist family = ScalarDist::Gaussian, Axis major_axis = Axis::Long)
        : n_rows(n_rows),
          n_cols(n_cols),
          major_axis(major_axis),
          dim_major((major_axis == Axis::Long) ? std::max(n_rows, n_cols) : std::min(n_rows, n_cols)),
          dim_minor((major_axis == Axis::Long) ? std::min(n_rows, n_cols) : std::max(n_rows, n_cols)),
          isometry_scale(std::pow((double) dim_minor, -0.5)),
          family(family),
          natural_layout(dense::natural_layout(major_axis, n_rows, n_cols)) {
        randblas_require(n_rows > 0);
        randblas_require(n_cols > 0);
    }

    template <typename T, typename RNG = DefaultRNG>
    DenseSkOp<T, RNG> sample(RNGState<RNG>& seed_state) {
        return {*this, seed_state};
    }
};

namespace dense {
// dense_skops.hh:172-185 -- pure integer arithmetic, done by the library so that both sides agree
template <typename RNG>
inline RNGState<RNG> compute_next_state(const DenseDist& D, const RNGState<RNG>& state) {
    uint32_t next[4];
    internal::check(rb_dense_next_state(D.n_rows, D.n_cols, (char) D.family, (char) D.major_axis, state.counter.v, next),
                    __func__);
    return internal::with_counter(state, next);
}
}  // namespace dense

// dense_skops.hh:357-478
template <typename T, typename RNG = DefaultRNG>
struct DenseSkOp {
    using distribution_t = DenseDist;
    using state_t = RNGState<RNG>;
    using scalar_t = T;

    const DenseDist dist;
    const state_t seed_state;
    const state_t next_state;
    const int64_t n_rows;
    const int64_t n_cols;
    bool own_memory;
    T* buff;
    const blas::Layout layout;

    DenseSkOp(DenseDist dist, const state_t& seed_state)
        : dist(dist),
          seed_state(seed_state),
          next_state(dense::compute_next_state(dist, seed_state)),
          n_rows(dist.n_rows),
          n_cols(dist.n_cols),
          own_memory(true),
          buff(nullptr),
          layout(dist.natural_layout) {}

    DenseSkOp(DenseSkOp<T, RNG>&& S)
        : dist(S.dist),
          seed_state(S.seed_state),
          next_state(S.next_state),
          n_rows(dist.n_rows),
          n_cols(dist.n_cols),
          own_memory(S.own_memory),
          buff(S.buff),
          layout(S.layout) {
        S.buff = nullptr;
    }

    ~DenseSkOp() {
        if (own_memory && buff != nullptr) delete[] buff;
    }
};

namespace internal {
inline int fill_dense_c(char layout, const DenseDist& D, int64_t n_rows, int64_t n_cols, int64_t ro_s, int64_t co_s,
                        float* buff, const uint32_t* ctr, const uint32_t* key, uint32_t* next) {
    return rb_fill_dense_f32(layout, D.n_rows, D.n_cols, (char) D.family, (char) D.major_axis, n_rows, n_cols, ro_s, co_s,
                             buff, 0, ctr, key, next, nullptr);
}
inline int fill_dense_c(char layout, const DenseDist& D, int64_t n_rows, int64_t n_cols, int64_t ro_s, int64_t co_s,
                        double* buff, const uint32_t* ctr, const uint32_t* key, uint32_t* next) {
    return rb_fill_dense_f64(layout, D.n_rows, D.n_cols, (char) D.family, (char) D.major_axis, n_rows, n_cols, ro_s, co_s,
                             buff, 0, ctr, key, next, nullptr);
}
}  // namespace internal

// dense_skops.hh:563-606. `buff` may be host memory (as with the reference) or device memory.
template <typename T, typename RNG = DefaultRNG>
RNGState<RNG> fill_dense_unpacked(blas::Layout layout, const DenseDist& D, int64_t n_rows, int64_t n_cols, int64_t ro_s,
                                  int64_t co_s, T* buff, const RNGState<RNG>& seed) {
    randblas_require(D.n_rows >= n_rows + ro_s);
    randblas_require(D.n_cols >= n_cols + co_s);
    uint32_t next[4];
    internal::check(internal::fill_dense_c(internal::to_char(layout), D, n_rows, n_cols, ro_s, co_s, buff, seed.counter.v,
                                           seed.key.v, next),
                    __func__);
    return internal::with_counter(seed, next);
}

// dense_skops.hh:623-626
template <typename T, typename RNG = DefaultRNG>
RNGState<RNG> fill_dense(const DenseDist& D, T* buff, const RNGState<RNG>& seed) {
    return fill_dense_unpacked(D.natural_layout, D, D.n_rows, D.n_cols, 0, 0, buff, seed);
}

// dense_skops.hh:649-658: allocates S.buff (host memory, new[]) if the operator owns its memory and has none
template <typename DenseSkOp>
void fill_dense(DenseSkOp& S) {
    if (S.own_memory && S.buff == nullptr) {
        using T = typename DenseSkOp::scalar_t;
        S.buff = new T[S.n_rows * S.n_cols];
    }
    randblas_require(S.buff != nullptr);
    fill_dense_unpacked(S.layout, S.dist, S.n_rows, S.n_cols, 0, 0, S.buff, S.seed_state);
}

}  // namespace RandBLAS
