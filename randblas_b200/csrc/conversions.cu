// Device-side format conversions for sparse DATA matrices: the step in front of sketch_sparse / left_spmm when the
// caller's matrix is not in the format it wants.
//
// Replaces (reference file:line): coo_to_csr / coo_to_csc (RandBLAS/sparse_data/conversions.hh:79-121: deep copy,
// COOMatrix::sort_arrays -- a std::sort of index pairs, coo_matrix.hh -- then sorted_idxs_to_compressed_ptr,
// sparse_data/base.hh:279-301) and csr_to_coo / csc_to_coo (conversions.hh:49-75).
//
// coo_to_compressed: one 64-bit key per nonzero (major * n_minor + minor), cub radix sort of (key, position) pairs over
// the significant bits only, a gather of the three arrays, and the compressed pointer from a histogram + scan.
// Entries come out ordered by (major, minor); entries with equal coordinates keep their input order (the radix sort
// is stable; std::sort in the reference leaves that order unspecified).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "common.cuh"
#include "kernels.h"

namespace rb {

namespace {

template <typename IDX>
__global__ void __launch_bounds__(256) coo_keys_kernel(int64_t nnz, const IDX* __restrict__ major, const IDX* __restrict__ minor,
                                                       int64_t n_minor, unsigned long long* __restrict__ keys,
                                                       unsigned int* __restrict__ pos) {
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t) gridDim.x * blockDim.x) {
        keys[e] = (unsigned long long) major[e] * (unsigned long long) n_minor + (unsigned long long) minor[e];
        pos[e] = (unsigned int) e;
    }
}

template <typename IDX, typename VAL>
__global__ void __launch_bounds__(256) coo_gather_kernel(int64_t nnz, const unsigned long long* __restrict__ keys,
                                                         const unsigned int* __restrict__ pos, int64_t n_minor,
                                                         const VAL* __restrict__ vals, VAL* __restrict__ ovals,
                                                         IDX* __restrict__ ominor, IDX* __restrict__ omajor,
                                                         unsigned long long* __restrict__ counts, unsigned long long n_major) {
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t) gridDim.x * blockDim.x) {
        const unsigned long long k = keys[e];
        unsigned long long mj = k / (unsigned long long) n_minor;
        if (mj >= n_major) mj = n_major - 1;      // out-of-range input index (invalid matrix): never write outside counts[]
        ominor[e] = (IDX) (k - mj * (unsigned long long) n_minor);
        if (omajor) omajor[e] = (IDX) mj;
        ovals[e] = vals[pos[e]];
        atomicAdd(counts + mj, 1ull);
    }
}

template <typename IDX>
__global__ void __launch_bounds__(256) store_ptr_kernel(int64_t n, const unsigned long long* __restrict__ scan, IDX* __restrict__ ptr) {
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
        ptr[i] = (IDX) scan[i];
}

template <typename IDX, typename VAL>
int coo_to_compressed_t(int64_t n_major, int64_t n_minor, int64_t nnz, const void* vals, const void* major, const void* minor,
                        void* ovals, void* ominor, void* optr, void* omajor, cudaStream_t st) {
    const int64_t cap = (int64_t) sm_count() * 8;
    int64_t grid = (nnz + 255) / 256;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    // the sort key is major * n_minor + minor in 64 bits
    unsigned long long prod = 0;
    if (__builtin_mul_overflow((unsigned long long) (n_major > 0 ? n_major : 1), (unsigned long long) (n_minor > 0 ? n_minor : 1), &prod))
        return fail("coo_to_compressed: n_rows * n_cols must be below 2^64");
    unsigned long long* keys = (unsigned long long*) workspace(0, (size_t) (nnz + 1) * 8 * 2, st);
    unsigned int* pos = (unsigned int*) workspace(1, (size_t) (nnz + 1) * 4 * 2, st);
    unsigned long long* counts = (unsigned long long*) workspace(2, (size_t) (n_major + 2) * 8 * 2, st);
    if (!keys || !pos || !counts) return fail_cuda(cudaErrorMemoryAllocation, "conversion workspace");
    unsigned long long* keys2 = keys + nnz + 1;
    unsigned int* pos2 = pos + nnz + 1;
    unsigned long long* scan = counts + n_major + 2;
    RB_CUDA(cudaMemsetAsync(counts, 0, (size_t) (n_major + 1) * 8, st));
    if (nnz > 0) {
        coo_keys_kernel<IDX><<<(unsigned) grid, 256, 0, st>>>(nnz, (const IDX*) major, (const IDX*) minor, n_minor, keys, pos);
        count_launch();
        int end_bit = 1;
        {
            const unsigned long long span = (unsigned long long) n_major * (unsigned long long) n_minor;
            while (end_bit < 64 && (span >> end_bit) != 0) ++end_bit;
        }
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, pos, pos2, (int) nnz, 0, end_bit, st);
        void* tmp = workspace(3, tmp_bytes, st);
        if (!tmp) return fail_cuda(cudaErrorMemoryAllocation, "sort workspace");
        RB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, pos, pos2, (int) nnz, 0, end_bit, st));
        count_launch();
        coo_gather_kernel<IDX, VAL><<<(unsigned) grid, 256, 0, st>>>(nnz, keys2, pos2, n_minor, (const VAL*) vals, (VAL*) ovals,
                                                                    (IDX*) ominor, (IDX*) omajor, counts,
                                                                    (unsigned long long) n_major);
        count_launch();
    }
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, scan, (int) (n_major + 1), st);
    void* tmp = workspace(4, tmp_bytes, st);
    if (!tmp) return fail_cuda(cudaErrorMemoryAllocation, "scan workspace");
    RB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, scan, (int) (n_major + 1), st));
    count_launch();
    int64_t g2 = (n_major + 1 + 255) / 256;
    if (g2 > cap) g2 = cap;
    store_ptr_kernel<IDX><<<(unsigned) g2, 256, 0, st>>>(n_major + 1, scan, (IDX*) optr);
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

// All pointers are device pointers. major/minor: the COO index arrays along the compressed / the other axis
// (rows, cols for CSR). optr gets n_major + 1 entries; omajor (optional) the sorted major indices.
int launch_coo_to_compressed(int64_t n_major, int64_t n_minor, int64_t nnz, const void* vals, int val_bytes, const void* major,
                             const void* minor, int idx_bytes, void* ovals, void* ominor, void* optr, void* omajor,
                             cudaStream_t st) {
    if (nnz >= 0x7fffffffLL || n_major >= 0x7fffffffLL) return fail("conversion: more than 2^31 - 1 nonzeros or major indices");
    if (idx_bytes == 4) {
        if (val_bytes == 4) return coo_to_compressed_t<int32_t, float>(n_major, n_minor, nnz, vals, major, minor, ovals, ominor, optr, omajor, st);
        return coo_to_compressed_t<int32_t, double>(n_major, n_minor, nnz, vals, major, minor, ovals, ominor, optr, omajor, st);
    }
    if (val_bytes == 4) return coo_to_compressed_t<int64_t, float>(n_major, n_minor, nnz, vals, major, minor, ovals, ominor, optr, omajor, st);
    return coo_to_compressed_t<int64_t, double>(n_major, n_minor, nnz, vals, major, minor, ovals, ominor, optr, omajor, st);
}

}  // namespace rb
