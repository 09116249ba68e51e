// Internal launcher declarations shared by the translation units of librandblas_b200.so.
// All pointers are DEVICE pointers here; host staging lives in c_abi.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"

namespace rb {

void count_launch(int n = 1);
void count_tc_launch();
void count_owner_launch();

int launch_philox_words(Ctr128 ctr, PhiloxKey key, int64_t n_blocks, uint32_t* out, cudaStream_t st);

int launch_boxmuller_words(int64_t n, const uint32_t* w0, const uint32_t* w1, float* g0, float* g1, cudaStream_t st);

template <typename T>
int launch_fill_dense(const DenseGen& g, char family, int64_t v0, int64_t nv, int64_t u0, int64_t nu, T* dst,
                      int64_t sv, int64_t su, cudaStream_t st);

// SASO generation. Any of idxs_minor / vals may be null. idx_bytes, val_bytes in {4, 8}.
int launch_saso(Ctr128 ctr, PhiloxKey key, int64_t vec_nnz, int64_t dim_major, int64_t dim_minor, void* idxs_major,
                void* idxs_minor, int idx_bytes, void* vals, int val_bytes, cudaStream_t st);

// LASO generation (laso.cu): compacted COO arrays in first-occurrence order, entry count to *nnz_host (synchronises)
int launch_laso(Ctr128 ctr, PhiloxKey key, int64_t vec_nnz, int64_t dim_major, int64_t dim_minor, void* idxs_long,
                void* idxs_short, int idx_bytes, void* vals, int val_bytes, int64_t* nnz_host, cudaStream_t st);

// Index-sampling utilities (sampling.cu): RandBLAS/util.hh:459-560. rademachers may be null.
int launch_sample_indices_iid_uniform(Ctr128 ctr, PhiloxKey key, int64_t n, int64_t k, void* samples, int idx_bytes,
                                      void* rademachers, int val_bytes, cudaStream_t st);
int launch_sample_indices_iid(Ctr128 ctr, PhiloxKey key, int64_t n, const void* cdf, int val_bytes, int64_t k,
                              void* samples, int idx_bytes, cudaStream_t st);
template <typename T>
int launch_weights_to_cdf(int64_t n, T* w, T error_if_below, cudaStream_t st);   // synchronises

// Canonical dense problem: C(P x Q) = alpha * X(P x K) * Y(K x Q) + beta * C where X = op(S window) and
// Y, C are strided views (element (k,j) of Y at Y + k*yrs + j*ycs; element (i,j) of C at C + i*crs + j*ccs).
// Right sketches are mapped to this form by transposition in the C-ABI layer.
template <typename T>
struct DenseProblem {
    int64_t P, Q, K;
    T alpha, beta;
    // operator window: element (i,k) of X is the operator entry (row0 + i*rsi + k*rsk ... ) expressed as
    // natural coordinates: v = v0 + i*vi + k*vk, u = u0 + i*ui + k*uk, with exactly one of (vi,ui) and one
    // of (vk,uk) equal to 1 and the other 0.
    DenseGen gen;
    char family;
    int64_t v0, u0;
    int vi, ui, vk, uk;
    // materialised operator (S.buff) in natural layout with leading dimension dim_major, or null
    const T* S_buff;
    int64_t S_ld;
    const T* Y;
    int64_t yrs, ycs;
    T* C;
    int64_t crs, ccs;
};

template <typename T>
int launch_dense_generic(const DenseProblem<T>& p, cudaStream_t st);

// Tensor-core fast paths. Return 0 if launched, -1 if the shape/layout is not supported (caller falls
// back to the generic kernel), >0 on error.
int launch_dense_tc_f32(const DenseProblem<float>& p, cudaStream_t st);
int launch_dense_dmma_f64(const DenseProblem<double>& p, cudaStream_t st);

// SASO operator (wide or tall, Short axis) applied to strided dense data, canonical left form:
// C(P x Q) = alpha * X(P x K) * Y(K x Q) + beta * C, X = op(S_sparse window).
template <typename T>
struct SasoProblem {
    int64_t P, Q, K;
    T alpha, beta;
    Ctr128 ctr;
    PhiloxKey key;
    int64_t vec_nnz, dim_major, dim_minor;
    int major_is_rows;   // 1: the short (major) axis indexes rows of S (wide operator), 0: columns (tall)
    int x_is_transposed; // 1: X = S^T window
    int64_t ro_s, co_s;  // window origin in S coordinates; window is rs x cs
    int64_t rs, cs;
    const T* Y;
    int64_t yrs, ycs;
    T* C;
    int64_t crs, ccs;
};
template <typename T>
int launch_saso_apply(const SasoProblem<T>& p, cudaStream_t st);
// fast path (saso_binned.cu): one-time binning pre-pass + decoupled TMA pipeline. 0 = done, -1 = shape/layout not taken
// (the caller uses the atomic kernel), >0 error. C already beta-scaled.
template <typename T>
int launch_saso_binned(const SasoProblem<T>& p, cudaStream_t st);

// generic COO x dense: C(P x Q) += alpha * X * Y with X given by triplets inside a window
template <typename T>
struct CooProblem {
    int64_t P, Q, K;
    T alpha, beta;
    int64_t nnz;
    const T* vals;
    const void* rows;
    const void* cols;
    int idx_bytes;
    int x_is_transposed;
    int64_t ro_s, co_s, rs, cs;
    const T* Y;
    int64_t yrs, ycs;
    T* C;
    int64_t crs, ccs;
};
template <typename T>
int launch_coo_apply(const CooProblem<T>& p, cudaStream_t st);

// CSR rowptr / CSC colptr expanded to one index per stored entry (device pointers)
int launch_expand_ptr(int64_t n_major, const void* ptr, void* out, int idx_bytes, cudaStream_t st);

// COO -> CSR / CSC on the device (conversions.cu); device pointers
int launch_coo_to_compressed(int64_t n_major, int64_t n_minor, int64_t nnz, const void* vals, int val_bytes, const void* major,
                             const void* minor, int idx_bytes, void* ovals, void* ominor, void* optr, void* omajor,
                             cudaStream_t st);

// symmetry check of an n x n strided matrix (util.hh:128-148); *first_dev = i * n + j of the first violation or ~0
template <typename T>
int launch_symmetry_check(const T* A, int64_t n, int64_t rs, int64_t cs, T tol, unsigned long long* first_dev, cudaStream_t st);

// beta pre-scale of a strided P x Q matrix (beta == 0 writes zeros without reading)
template <typename T>
int launch_scale(int64_t P, int64_t Q, T beta, T* C, int64_t crs, int64_t ccs, cudaStream_t st);

// dense operator x sparse data, canonical left form C(P x Q) = alpha * X(P x K) * Ysp(K x Q) + beta * C with
// X = op(S window) generated on the fly and Ysp = op(A_sp window).
template <typename T>
struct SpDataProblem {
    int64_t P, Q, K;
    T alpha, beta;
    DenseGen gen;
    char family;
    int64_t v0, u0;
    int vi, ui, vk, uk;
    int fmt;                 // 0 CSR, 1 CSC, 2 COO  (of the stored matrix A_sp)
    int y_is_transposed;     // 1: Ysp = A_sp^T window
    int64_t A_rows, A_cols, nnz;
    const T* vals;
    const void* idx0;
    const void* idx1;
    int idx_bytes;
    int64_t ro_a, co_a;
    T* C;
    int64_t crs, ccs;
};
template <typename T>
int launch_spdata(const SpDataProblem<T>& p, cudaStream_t st);

// cached device workspace (grow-only), slot in [0, RB_WS_SLOTS), owned by (current device, stream, calling thread)
#define RB_WS_SLOTS 10
void* workspace(int slot, size_t bytes, cudaStream_t st);
void release_workspace();

int64_t get_option(const char* name);

}  // namespace rb
