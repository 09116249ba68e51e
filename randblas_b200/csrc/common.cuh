// Shared host/device helpers for the C-ABI translation units.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>
#include <cuda_runtime.h>
#include "philox.cuh"

namespace rb {

// thread-local error string behind rb_last_error()
void set_error(const std::string& msg);
int fail(const std::string& msg);          // sets the error, returns RB_ERR_ARG (1)
int fail_cuda(cudaError_t e, const char* what);  // returns RB_ERR_CUDA (2)

#define RB_REQUIRE(cond)                                                                          \
    do {                                                                                          \
        if (!(cond)) return ::rb::fail(std::string("(" #cond ") was required, but did not hold, in function ") + __func__); \
    } while (0)

#define RB_CUDA(call)                                                   \
    do {                                                                \
        cudaError_t e_ = (call);                                        \
        if (e_ != cudaSuccess) return ::rb::fail_cuda(e_, #call);       \
    } while (0)

// DenseDist as the reference derives it (RandBLAS/dense_skops.hh:187-199, 300-328)
struct DenseDistInfo {
    int64_t n_rows, n_cols, dim_major, dim_minor;
    char family, major_axis, natural_layout;
    bool ok;
};
inline DenseDistInfo make_dense_dist(int64_t n_rows, int64_t n_cols, char family, char major_axis) {
    DenseDistInfo D{};
    D.ok = n_rows > 0 && n_cols > 0 && (family == 'G' || family == 'U') && (major_axis == 'S' || major_axis == 'L');
    if (!D.ok) return D;
    int64_t mx = n_rows > n_cols ? n_rows : n_cols, mn = n_rows < n_cols ? n_rows : n_cols;
    D.n_rows = n_rows; D.n_cols = n_cols; D.family = family; D.major_axis = major_axis;
    D.dim_major = (major_axis == 'L') ? mx : mn;
    D.dim_minor = (major_axis == 'L') ? mn : mx;
    bool wide = n_rows < n_cols, lng = major_axis == 'L';
    D.natural_layout = (wide && lng) ? 'R' : (wide ? 'C' : (lng ? 'C' : 'R'));
    return D;
}

// SparseDist (RandBLAS/sparse_skops.hh:205-224)
struct SparseDistInfo {
    int64_t n_rows, n_cols, dim_major, dim_minor, vec_nnz, full_nnz;
    char major_axis;
    bool ok;
};
inline SparseDistInfo make_sparse_dist(int64_t n_rows, int64_t n_cols, int64_t vec_nnz, char major_axis) {
    SparseDistInfo D{};
    D.ok = n_rows > 0 && n_cols > 0 && vec_nnz > 0 && (major_axis == 'S' || major_axis == 'L');
    if (!D.ok) return D;
    int64_t mx = n_rows > n_cols ? n_rows : n_cols, mn = n_rows < n_cols ? n_rows : n_cols;
    D.n_rows = n_rows; D.n_cols = n_cols; D.vec_nnz = vec_nnz; D.major_axis = major_axis;
    D.dim_major = (major_axis == 'S') ? mn : mx;
    D.dim_minor = n_rows + n_cols - D.dim_major;
    D.full_nnz = vec_nnz * D.dim_minor;
    D.ok = vec_nnz <= D.dim_major;
    return D;
}

inline Ctr128 load_ctr(const uint32_t* c) { return Ctr128{c[0], c[1], c[2], c[3]}; }
inline void store_ctr(Ctr128 c, uint32_t* out) {
    if (out) { out[0] = c.c0; out[1] = c.c1; out[2] = c.c2; out[3] = c.c3; }
}

int sm_count();   // SMs of the current device (cached)

// One-time initialisation flag PER DEVICE (function attributes such as the dynamic shared-memory opt-in belong to a
// device's context, not to the process): need() is true until done() was called on the current device.
struct DevOnce {
    std::atomic<uint64_t> mask{0};
    static int dev() { int d = 0; if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) { cudaGetLastError(); return -1; } return d; }
    bool need() const { const int d = dev(); return d < 0 || !(mask.load(std::memory_order_acquire) & (1ull << d)); }
    void done() { const int d = dev(); if (d >= 0) mask.fetch_or(1ull << d, std::memory_order_release); }
};
// device copy of the folded logf table (philox.cuh), built once per device; nullptr on failure
const double2* logf_table_device();

// How a DenseSkOp window is addressed by the fused kernels: element (i, j) of the operator
// (absolute row/col indices into the full D.n_rows x D.n_cols sample).
struct DenseGen {
    Ctr128 ctr;        // seed counter
    PhiloxKey key;
    int64_t R;         // Philox blocks per major-axis vector = ceil(dim_major / 4)
    int nat_row;       // 1: natural layout RowMajor => (v,u) = (row,col); 0: (v,u) = (col,row)
    const double2* logtab;   // folded logf table in global memory (Gaussian family)
};
inline DenseGen make_dense_gen(const DenseDistInfo& D, const uint32_t* ctr, const uint32_t* key) {
    DenseGen g;
    g.ctr = load_ctr(ctr);
    g.key = PhiloxKey{key[0], key[1]};
    g.R = (D.dim_major + 3) / 4;
    g.nat_row = D.natural_layout == 'R';
    g.logtab = (D.family == 'G') ? logf_table_device() : nullptr;
    return g;
}

}  // namespace rb
