// randblas_b200 -- header-only drop-in layer, part 4: sparse matrix containers (views and owners).
// Field names and constructor signatures follow RandBLAS/sparse_data/{base,coo_matrix,csr_matrix,csc_matrix}.hh;
// conversions, sorting and the CPU SpMM kernels of the reference are outside the hot path.
#pragma once
#include <algorithm>
#include <type_traits>
#include "base.hh"

namespace RandBLAS {
namespace sparse_data {

enum class IndexBase : char { Zero = 'Z', One = 'O' };                          // sparse_data/base.hh:52
enum class NonzeroSort : char { CSC = 'C', CSR = 'R', None = 'N' };             // sparse_data/base.hh:108-118

// csr_matrix.hh:73-245
template <typename T, typename sint_t = int64_t>
struct CSRMatrix {
    using scalar_t = T;
    using index_t = sint_t;
    const int64_t n_rows;
    const int64_t n_cols;
    bool own_memory;
    int64_t nnz;
    IndexBase index_base;
    T* vals;
    sint_t* rowptr;
    sint_t* colidxs;
    CSRMatrix(int64_t n_rows, int64_t n_cols)
        : n_rows(n_rows), n_cols(n_cols), own_memory(true), nnz(0), index_base(IndexBase::Zero), vals(nullptr),
          rowptr(nullptr), colidxs(nullptr) {}
    CSRMatrix(int64_t n_rows, int64_t n_cols, int64_t nnz, T* vals, sint_t* rowptr, sint_t* colidxs,
              IndexBase index_base = IndexBase::Zero)
        : n_rows(n_rows), n_cols(n_cols), own_memory(false), nnz(nnz), index_base(index_base), vals(vals), rowptr(rowptr),
          colidxs(colidxs) {}
    // move constructor (csr_matrix.hh:170-176): the source gives up its arrays
    CSRMatrix(CSRMatrix<T, sint_t>&& o)
        : n_rows(o.n_rows), n_cols(o.n_cols), own_memory(o.own_memory), nnz(o.nnz), index_base(o.index_base), vals(o.vals),
          rowptr(o.rowptr), colidxs(o.colidxs) {
        o.vals = nullptr; o.rowptr = nullptr; o.colidxs = nullptr; o.nnz = 0;
    }
    ~CSRMatrix() {
        if (own_memory) { delete[] vals; delete[] rowptr; delete[] colidxs; }
    }
    void reserve(int64_t arg_nnz) {                                              // csr_matrix.hh:186-197
        randblas_require(arg_nnz > 0 && own_memory && colidxs == nullptr && vals == nullptr);
        if (rowptr == nullptr) rowptr = new sint_t[n_rows + 1]{};
        nnz = arg_nnz;
        colidxs = new sint_t[nnz]{};
        vals = new T[nnz]{};
    }
};

// csc_matrix.hh:74-244
template <typename T, typename sint_t = int64_t>
struct CSCMatrix {
    using scalar_t = T;
    using index_t = sint_t;
    const int64_t n_rows;
    const int64_t n_cols;
    bool own_memory;
    int64_t nnz;
    IndexBase index_base;
    T* vals;
    sint_t* rowidxs;
    sint_t* colptr;
    CSCMatrix(int64_t n_rows, int64_t n_cols)
        : n_rows(n_rows), n_cols(n_cols), own_memory(true), nnz(0), index_base(IndexBase::Zero), vals(nullptr),
          rowidxs(nullptr), colptr(nullptr) {}
    CSCMatrix(int64_t n_rows, int64_t n_cols, int64_t nnz, T* vals, sint_t* rowidxs, sint_t* colptr,
              IndexBase index_base = IndexBase::Zero)
        : n_rows(n_rows), n_cols(n_cols), own_memory(false), nnz(nnz), index_base(index_base), vals(vals),
          rowidxs(rowidxs), colptr(colptr) {}
    CSCMatrix(CSCMatrix<T, sint_t>&& o)                                           // csc_matrix.hh:169-175
        : n_rows(o.n_rows), n_cols(o.n_cols), own_memory(o.own_memory), nnz(o.nnz), index_base(o.index_base), vals(o.vals),
          rowidxs(o.rowidxs), colptr(o.colptr) {
        o.vals = nullptr; o.rowidxs = nullptr; o.colptr = nullptr; o.nnz = 0;
    }
    ~CSCMatrix() {
        if (own_memory) { delete[] vals; delete[] rowidxs; delete[] colptr; }
    }
    void reserve(int64_t arg_nnz) {
        randblas_require(arg_nnz > 0 && own_memory && rowidxs == nullptr && vals == nullptr);
        if (colptr == nullptr) colptr = new sint_t[n_cols + 1]{};
        nnz = arg_nnz;
        rowidxs = new sint_t[nnz]{};
        vals = new T[nnz]{};
    }
};

// coo_matrix.hh:82-308
template <typename T, typename sint_t = int64_t>
struct COOMatrix {
    using scalar_t = T;
    using index_t = sint_t;
    const int64_t n_rows;
    const int64_t n_cols;
    bool own_memory;
    int64_t nnz;
    IndexBase index_base;
    T* vals;
    sint_t* rows;
    sint_t* cols;
    NonzeroSort sort;
    COOMatrix(int64_t n_rows, int64_t n_cols)
        : n_rows(n_rows), n_cols(n_cols), own_memory(true), nnz(0), index_base(IndexBase::Zero), vals(nullptr),
          rows(nullptr), cols(nullptr), sort(NonzeroSort::None) {}
    // view constructor. The reference scans the arrays here to classify their sort order (coo_matrix.hh:188-189);
    // the GPU kernels do not depend on the order, so the tag is left at None unless the caller sets it.
    COOMatrix(int64_t n_rows, int64_t n_cols, int64_t nnz, T* vals, sint_t* rows, sint_t* cols, bool compute_sort_type = true,
              IndexBase index_base = IndexBase::Zero)
        : n_rows(n_rows), n_cols(n_cols), own_memory(false), nnz(nnz), index_base(index_base), vals(vals), rows(rows),
          cols(cols), sort(NonzeroSort::None) {
        (void) compute_sort_type;
    }
    COOMatrix(COOMatrix<T, sint_t>&& o)                                           // coo_matrix.hh:195-202
        : n_rows(o.n_rows), n_cols(o.n_cols), own_memory(o.own_memory), nnz(o.nnz), index_base(o.index_base), vals(o.vals),
          rows(o.rows), cols(o.cols), sort(o.sort) {
        o.vals = nullptr; o.rows = nullptr; o.cols = nullptr; o.nnz = 0;
    }
    ~COOMatrix() {
        if (own_memory) { delete[] vals; delete[] rows; delete[] cols; }
    }
    void reserve(int64_t arg_nnz) {
        randblas_require(arg_nnz > 0 && own_memory && vals == nullptr && rows == nullptr && cols == nullptr);
        nnz = arg_nnz;
        vals = new T[nnz]{};
        rows = new sint_t[nnz]{};
        cols = new sint_t[nnz]{};
    }
};

}  // namespace sparse_data

using sparse_data::COOMatrix;
using sparse_data::CSCMatrix;
using sparse_data::CSRMatrix;
using sparse_data::IndexBase;
using sparse_data::NonzeroSort;

namespace internal {
// (fmt, idx0, idx1) as the C ABI wants them (include/randblas_b200.h, K4)
template <typename T, typename I>
inline void sp_arrays(const CSRMatrix<T, I>& A, int& fmt, const void*& i0, const void*& i1) { fmt = 0; i0 = A.rowptr; i1 = A.colidxs; }
template <typename T, typename I>
inline void sp_arrays(const CSCMatrix<T, I>& A, int& fmt, const void*& i0, const void*& i1) { fmt = 1; i0 = A.rowidxs; i1 = A.colptr; }
template <typename T, typename I>
inline void sp_arrays(const COOMatrix<T, I>& A, int& fmt, const void*& i0, const void*& i1) { fmt = 2; i0 = A.rows; i1 = A.cols; }
}  // namespace internal

namespace sparse_data {
// conversions.hh:101-121 / :79-99. The destination must be an empty owning matrix of the same shape (reserve() is
// called here, as in the reference); the arrays are host memory, the sort runs on the device (rb_coo_to_compressed).
template <typename T, typename I1, typename I2>
void coo_to_csr(const COOMatrix<T, I1>& coo, CSRMatrix<T, I2>& csr) {
    static_assert(std::is_same_v<I1, I2>, "index types must match in this build");
    randblas_require(csr.n_rows == coo.n_rows);
    randblas_require(csr.n_cols == coo.n_cols);
    randblas_require(csr.index_base == IndexBase::Zero);
    if (coo.nnz == 0) return;
    csr.reserve(coo.nnz);
    internal::check(rb_coo_to_compressed(0, coo.n_rows, coo.n_cols, coo.nnz, coo.vals, (int) sizeof(T), coo.rows, coo.cols,
                                         (int) sizeof(I1), csr.vals, csr.colidxs, csr.rowptr, nullptr),
                    __func__);
}
template <typename T, typename I1, typename I2>
void coo_to_csc(const COOMatrix<T, I1>& coo, CSCMatrix<T, I2>& csc) {
    static_assert(std::is_same_v<I1, I2>, "index types must match in this build");
    randblas_require(csc.n_rows == coo.n_rows);
    randblas_require(csc.n_cols == coo.n_cols);
    randblas_require(csc.index_base == IndexBase::Zero);
    if (coo.nnz == 0) return;
    csc.reserve(coo.nnz);
    internal::check(rb_coo_to_compressed(1, coo.n_rows, coo.n_cols, coo.nnz, coo.vals, (int) sizeof(T), coo.rows, coo.cols,
                                         (int) sizeof(I1), csc.vals, csc.rowidxs, csc.colptr, nullptr),
                    __func__);
}
// conversions.hh:63-75 / :49-61
template <typename T, typename I1, typename I2>
void csr_to_coo(const CSRMatrix<T, I1>& csr, COOMatrix<T, I2>& coo) {
    static_assert(std::is_same_v<I1, I2>, "index types must match in this build");
    randblas_require(csr.n_rows == coo.n_rows);
    randblas_require(csr.n_cols == coo.n_cols);
    if (csr.nnz == 0) return;
    coo.reserve(csr.nnz);
    internal::check(rb_expand_ptr(csr.n_rows, csr.rowptr, csr.nnz, coo.rows, (int) sizeof(I1), nullptr), __func__);
    std::copy(csr.vals, csr.vals + csr.nnz, coo.vals);
    std::copy(csr.colidxs, csr.colidxs + csr.nnz, coo.cols);
    coo.sort = NonzeroSort::CSR;
}
template <typename T, typename I1, typename I2>
void csc_to_coo(const CSCMatrix<T, I1>& csc, COOMatrix<T, I2>& coo) {
    static_assert(std::is_same_v<I1, I2>, "index types must match in this build");
    randblas_require(csc.n_rows == coo.n_rows);
    randblas_require(csc.n_cols == coo.n_cols);
    if (csc.nnz == 0) return;
    coo.reserve(csc.nnz);
    internal::check(rb_expand_ptr(csc.n_cols, csc.colptr, csc.nnz, coo.cols, (int) sizeof(I1), nullptr), __func__);
    std::copy(csc.vals, csc.vals + csc.nnz, coo.vals);
    std::copy(csc.rowidxs, csc.rowidxs + csc.nnz, coo.rows);
    coo.sort = NonzeroSort::CSC;
}
}  // namespace sparse_data

}  // namespace RandBLAS
