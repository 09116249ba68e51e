// randblas_b200 -- index-sampling utilities of the reference's RandBLAS/util.hh:440-560 (weights_to_cdf,
// sample_indices_iid, sample_indices_iid_uniform), same names, template parameters and argument order; each forwards
// to one C-ABI entry point (sampling.cu). Pointers may be host or device memory.
#pragma once
#include <cmath>
#include <limits>
#include <type_traits>
#include "base.hh"

namespace RandBLAS {

// util.hh:440-443
template <typename T>
T sqrt_epsilon() {
    return std::sqrt(std::numeric_limits<T>::epsilon());
}

// util.hh:459-473: w <- normalised running sum of max(w, 0); throws if a weight is below error_if_below or the total is
// below sqrt(n) * eps.
template <typename T>
void weights_to_cdf(int64_t n, T* w, T error_if_below = -sqrt_epsilon<T>()) {
    static_assert(std::is_same_v<T, float> || std::is_same_v<T, double>, "float or double");
    if constexpr (std::is_same_v<T, float>) internal::check(rb_weights_to_cdf_f32(n, w, error_if_below, nullptr), __func__);
    else internal::check(rb_weights_to_cdf_f64(n, w, error_if_below, nullptr), __func__);
}

// util.hh:490-513: k independent samples from the CDF over {0..n-1}
template <typename T, typename sint_t, typename state_t = RNGState<DefaultRNG>>
state_t sample_indices_iid(int64_t n, const T* cdf, int64_t k, sint_t* samples, const state_t& state) {
    uint32_t next[4];
    internal::check(rb_sample_indices_iid(n, cdf, (int) sizeof(T), k, samples, (int) sizeof(sint_t), state.counter.v,
                                          state.key.v, next, nullptr),
                    __func__);
    return internal::with_counter(state, next);
}

// util.hh:515-547: k independent uniform samples from {0..n-1}, optionally with Rademacher signs
template <typename T, typename sint_t, bool WriteRademachers = true, typename state_t = RNGState<DefaultRNG>>
state_t sample_indices_iid_uniform(int64_t n, int64_t k, sint_t* samples, T* rademachers, const state_t& state) {
    uint32_t next[4];
    internal::check(rb_sample_indices_iid_uniform(n, k, samples, (int) sizeof(sint_t),
                                                  WriteRademachers ? (void*) rademachers : nullptr, (int) sizeof(T),
                                                  state.counter.v, state.key.v, next, nullptr),
                    __func__);
    return internal::with_counter(state, next);
}

// util.hh:556-559
template <typename sint_t = int64_t, typename state_t = RNGState<DefaultRNG>>
state_t sample_indices_iid_uniform(int64_t n, int64_t k, sint_t* samples, const state_t& state) {
    return sample_indices_iid_uniform<float, sint_t, false, state_t>(n, k, samples, (float*) nullptr, state);
}

}  // namespace RandBLAS
