// Index-sampling utilities next to the sparse operators:
//   * sample_indices_iid_uniform<T, sint_t, WriteRademachers>   RandBLAS/util.hh:515-560
//   * sample_indices_iid<T, sint_t>                              RandBLAS/util.hh:490-513
//   * weights_to_cdf<T>                                          RandBLAS/util.hh:459-473
//
// Reference semantics restated. Both samplers walk ONE Philox stream: block b = seed + b, transformed by uneg11
// to four floats. sample_indices_iid_uniform without Rademachers takes one float per sample (sample i = lane i % 4
// of block i / 4), with Rademachers two (index from lane 2(i % 2), sign from lane 2(i % 2) + 1 of block i / 2):
//     index = (sint_t) ((double) (sint_t) n * (((double) x + 1.0) / 2.0))      util.hh:532-533 (truncation)
//     sign  = x' >= 0 ? +1 : -1        (<=> (int32) word >= 0)                   util.hh:536
// sample_indices_iid maps u = ((T) x + 1) / 2, in T arithmetic, through std::lower_bound on the CDF (first position
// whose value is not less than u). The returned state is seed + ceil(k / samples per block). Samples are independent
// of each other, so one thread per sample replaces the serial loop.
//
// weights_to_cdf is a running sum in T in index order followed by blas::scal with (T) 1 / sum; a parallel scan would
// round differently and move CDF entries by an ulp (and with them samples that fall on a boundary), so the sum is
// kept serial: one CTA stages 4096 weights at a time in shared memory, one thread adds them in order, all threads
// write the prefix back. The reference throws at the first weight below error_if_below AFTER having overwritten the
// entries before it; the same prefix is left here, and the failure comes back as an argument error.
#include <cmath>
#include <limits>
#include "common.cuh"
#include "kernels.h"

namespace rb {

namespace {

template <typename IDX, typename VAL, bool RAD>
__global__ void __launch_bounds__(256) iid_uniform_kernel(Ctr128 ctr, PhiloxKey key, int64_t n, int64_t k,
                                                          IDX* __restrict__ samples, VAL* __restrict__ rad) {
    constexpr int SPB = RAD ? 2 : 4;                    // samples per Philox block
    const int64_t nblk = (k + SPB - 1) / SPB;
    const double dN = (double) (IDX) n;                 // "(sint_t) dN * random_unif01"
    for (int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += (int64_t) gridDim.x * blockDim.x) {
        const uint4 w = philox4x32_10(ctr_add(ctr, (uint64_t) b), key);
        const uint32_t wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < SPB; ++j) {
            const int64_t i = b * SPB + j;
            if (i >= k) break;
            const uint32_t wa = RAD ? wv[2 * j] : wv[j];
            const double u = __ddiv_rn(__dadd_rn((double) uneg11f(wa), 1.0), 2.0);    // uneg11_to_u01<double>
            samples[i] = (IDX) __dmul_rn(dN, u);
            if constexpr (RAD) rad[i] = ((int32_t) wv[2 * j + 1] >= 0) ? (VAL) 1 : (VAL) -1;
        }
    }
}

template <typename T, typename IDX>
__global__ void __launch_bounds__(256) iid_cdf_kernel(Ctr128 ctr, PhiloxKey key, int64_t n, const T* __restrict__ cdf,
                                                      int64_t k, IDX* __restrict__ samples) {
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (int64_t) gridDim.x * blockDim.x) {
        const uint4 w = philox4x32_10(ctr_add(ctr, (uint64_t) (i >> 2)), key);
        const uint32_t wa = (i & 3) == 0 ? w.x : (i & 3) == 1 ? w.y : (i & 3) == 2 ? w.z : w.w;
        T u;
        if constexpr (sizeof(T) == 4) u = __fdiv_rn(__fadd_rn(uneg11f(wa), 1.0f), 2.0f);
        else u = __ddiv_rn(__dadd_rn((double) uneg11f(wa), 1.0), 2.0);
        int64_t lo = 0, len = n;                        // std::lower_bound: first position with !(cdf[pos] < u)
        while (len > 0) {
            const int64_t half = len >> 1;
            if (cdf[lo + half] < u) { lo += half + 1; len -= half + 1; }
            else len = half;
        }
        samples[i] = (IDX) lo;
    }
}

constexpr int CDF_THREADS = 256;
constexpr int CDF_CHUNK = 4096;

// out[0] = final sum, out[1] = index of the first weight that failed "val >= error_if_below" (n if none)
template <typename T>
__global__ void __launch_bounds__(CDF_THREADS) cdf_prefix_kernel(int64_t n, T* __restrict__ w, T error_if_below,
                                                                 T* __restrict__ out_sum, int64_t* __restrict__ out_bad) {
    __shared__ T s[CDF_CHUNK];
    __shared__ int s_stop;
    T sum = (T) 0;
    int64_t bad = n;
    for (int64_t base = 0; base < n; base += CDF_CHUNK) {
        const int len = (int) min((int64_t) CDF_CHUNK, n - base);
        for (int i = threadIdx.x; i < len; i += CDF_THREADS) s[i] = w[base + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            int stop = len;
            for (int i = 0; i < len; ++i) {
                T val = s[i];
                if (!(val >= error_if_below)) { stop = i; bad = base + i; break; }     // randblas_require, util.hh:465
                val = (val < (T) 0) ? (T) 0 : val;                                      // std::max(val, 0)
                if constexpr (sizeof(T) == 4) sum = __fadd_rn(sum, val); else sum = __dadd_rn(sum, val);
                s[i] = sum;
            }
            s_stop = stop;
        }
        __syncthreads();
        const int stop = s_stop;
        for (int i = threadIdx.x; i < stop; i += CDF_THREADS) w[base + i] = s[i];
        __syncthreads();
        if (stop < len) break;
    }
    if (threadIdx.x == 0) { *out_sum = sum; *out_bad = bad; }
}

template <typename T>
__global__ void __launch_bounds__(256) scal_kernel(int64_t n, T alpha, T* __restrict__ w) {
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
        if constexpr (sizeof(T) == 4) w[i] = __fmul_rn(w[i], alpha); else w[i] = __dmul_rn(w[i], alpha);
    }
}

inline unsigned grid_for(int64_t items) {
    int64_t g = (items + 255) / 256;
    const int64_t cap = (int64_t) sm_count() * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned) g;
}

template <typename IDX>
int iid_uniform_t(Ctr128 ctr, PhiloxKey key, int64_t n, int64_t k, void* samples, void* rad, int val_bytes, cudaStream_t st) {
    if (!rad) {
        iid_uniform_kernel<IDX, float, false><<<grid_for((k + 3) / 4), 256, 0, st>>>(ctr, key, n, k, (IDX*) samples, nullptr);
    } else if (val_bytes == 4) {
        iid_uniform_kernel<IDX, float, true><<<grid_for((k + 1) / 2), 256, 0, st>>>(ctr, key, n, k, (IDX*) samples, (float*) rad);
    } else {
        iid_uniform_kernel<IDX, double, true><<<grid_for((k + 1) / 2), 256, 0, st>>>(ctr, key, n, k, (IDX*) samples, (double*) rad);
    }
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int launch_sample_indices_iid_uniform(Ctr128 ctr, PhiloxKey key, int64_t n, int64_t k, void* samples, int idx_bytes,
                                      void* rademachers, int val_bytes, cudaStream_t st) {
    if (k <= 0) return 0;
    if (idx_bytes == 4) return iid_uniform_t<int32_t>(ctr, key, n, k, samples, rademachers, val_bytes, st);
    return iid_uniform_t<int64_t>(ctr, key, n, k, samples, rademachers, val_bytes, st);
}

int launch_sample_indices_iid(Ctr128 ctr, PhiloxKey key, int64_t n, const void* cdf, int val_bytes, int64_t k,
                              void* samples, int idx_bytes, cudaStream_t st) {
    if (k <= 0) return 0;
    const unsigned g = grid_for(k);
    if (val_bytes == 4) {
        if (idx_bytes == 4) iid_cdf_kernel<float, int32_t><<<g, 256, 0, st>>>(ctr, key, n, (const float*) cdf, k, (int32_t*) samples);
        else iid_cdf_kernel<float, int64_t><<<g, 256, 0, st>>>(ctr, key, n, (const float*) cdf, k, (int64_t*) samples);
    } else {
        if (idx_bytes == 4) iid_cdf_kernel<double, int32_t><<<g, 256, 0, st>>>(ctr, key, n, (const double*) cdf, k, (int32_t*) samples);
        else iid_cdf_kernel<double, int64_t><<<g, 256, 0, st>>>(ctr, key, n, (const double*) cdf, k, (int64_t*) samples);
    }
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}

// Returns 0, or RB_ERR_ARG with the reference's message shape when one of its two randblas_require lines fails.
template <typename T>
int launch_weights_to_cdf(int64_t n, T* w, T error_if_below, cudaStream_t st) {
    if (n <= 0) {
        // the loop is empty and sum = 0 < sqrt(n) * eps fails only for n > 0; n == 0: 0 >= 0 holds, scal of nothing
        return n == 0 ? 0 : fail("(n >= 0) was required, but did not hold, in function weights_to_cdf");
    }
    char* ws = (char*) workspace(0, 64, st);
    if (!ws) return fail_cuda(cudaErrorMemoryAllocation, "weights_to_cdf workspace");
    T* d_sum = (T*) ws;
    int64_t* d_bad = (int64_t*) (ws + 16);
    cdf_prefix_kernel<T><<<1, CDF_THREADS, 0, st>>>(n, w, error_if_below, d_sum, d_bad);
    count_launch();
    RB_CUDA(cudaGetLastError());
    T sum;
    int64_t bad;
    RB_CUDA(cudaMemcpyAsync(&sum, d_sum, sizeof(T), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    if (bad < n) return fail("(val >= error_if_below) was required, but did not hold, in function weights_to_cdf");
    // util.hh:470: sum >= ((T) std::sqrt(n)) * eps
    if (!(sum >= ((T) std::sqrt((double) n)) * std::numeric_limits<T>::epsilon()))
        return fail("(sum >= ((T) std::sqrt(n)) * std::numeric_limits<T>::epsilon()) was required, but did not hold, in function weights_to_cdf");
    scal_kernel<T><<<grid_for(n), 256, 0, st>>>(n, ((T) 1.0) / sum, w);      // blas::scal(n, 1 / sum, w, 1), util.hh:471
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}
template int launch_weights_to_cdf<float>(int64_t, float*, float, cudaStream_t);
template int launch_weights_to_cdf<double>(int64_t, double*, double, cudaStream_t);

}  // namespace rb
