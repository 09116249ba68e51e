/* randblas_b200 -- C ABI of the B200-native sketching hot path.
 *
 * This is the drop-in boundary: the header-only C++ layer in include/RandBLAS/ (same names and
 * signatures as the reference's public API) forwards each hot call to exactly one symbol below,
 * and any other host language can bind the same symbols.
 *
 * Conventions
 *   - plain C types only; sizes are int64_t; enums are chars with BLAS++'s values:
 *       layout 'C' (ColMajor) / 'R' (RowMajor); op 'N' / 'T';
 *       family 'G' (Gaussian) / 'U' (Uniform on [-sqrt3, sqrt3]); major_axis 'S' (Short) / 'L' (Long).
 *   - an RNGState is passed as its two arrays: ctr[4] (128-bit little-endian counter) and key[2]
 *     (reference: RandBLAS/base.hh:64-164). States are never mutated; functions that advance the
 *     stream write the advanced counter to next_ctr[4] (may be NULL).
 *   - data pointers (buff, A, B, vals, rows, cols, ...) may be DEVICE pointers (used in place, the
 *     measured mode) or HOST pointers (staged through device memory inside the call: this is what
 *     a caller of the CPU reference has). Both are detected with cudaPointerGetAttributes.
 *   - every call is ordered on `stream` (a cudaStream_t passed as void*, NULL = default stream).
 *     Calls with host buffers return after the results are back in host memory; calls with device
 *     buffers return after enqueueing (synchronise the stream to observe the result).
 *   - return value: 0 = ok; RB_ERR_ARG = an argument check failed (the reference would have thrown
 *     RandBLAS::Error before touching data, exceptions.hh:57-95); RB_ERR_CUDA = a CUDA error.
 *     rb_last_error() returns the thread-local message. No exception crosses this boundary.
 *   - the library never allocates or frees caller-visible memory. Internal workspace is cached per
 *     (device, stream, calling host thread), so calls on different streams or from different threads
 *     never share scratch memory; rb_release_workspace() frees it (no call may be in flight).
 */
#ifndef RANDBLAS_B200_H
#define RANDBLAS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB_OK 0
#define RB_ERR_ARG 1
#define RB_ERR_CUDA 2

const char* rb_last_error(void);
int rb_version(void);                 /* 100 * major + minor */
int rb_release_workspace(void);
/* info[0] = SM count, info[1] = compute capability * 10 + minor, info[2] = total global memory (bytes) */
int rb_device_info(int64_t info[3]);
/* cudaStreamSynchronize(stream) (NULL = default stream): used by the synchronous drop-in header layer */
int rb_sync_stream(void* stream);

/* ---- RNG state arithmetic (host-side, pure integer; RandBLAS/base.hh:116-119, Random123 array.h incr) ---- */
void rb_rngstate_from_u64(uint64_t k, uint32_t ctr[4], uint32_t key[2]);
void rb_ctr_incr(uint32_t ctr[4], uint64_t n);
/* DenseDist / SparseDist derived fields (RandBLAS/dense_skops.hh:231-350, sparse_skops.hh:131-246).
 * dense info = {dim_major, dim_minor, natural_layout}; sparse info = {dim_major, dim_minor, full_nnz} */
int rb_dense_dist_info(int64_t n_rows, int64_t n_cols, char family, char major_axis, int64_t info[3],
                       double* isometry_scale);
int rb_sparse_dist_info(int64_t n_rows, int64_t n_cols, int64_t vec_nnz, char major_axis, int64_t info[3],
                        double* isometry_scale);
/* next_state of DenseSkOp / SparseSkOp constructors (dense_skops.hh:172-185, sparse_skops.hh:266-283) */
int rb_dense_next_state(int64_t n_rows, int64_t n_cols, char family, char major_axis, const uint32_t ctr[4],
                        uint32_t next_ctr[4]);
int rb_sparse_next_state(int64_t n_rows, int64_t n_cols, int64_t vec_nnz, char major_axis, const uint32_t ctr[4],
                         uint32_t next_ctr[4]);

/* ---- raw generator output (for bit-exact checks of the device generator) ----
 * out[4*i .. 4*i+3] = Philox4x32_10(ctr + i, key), i in [0, n_blocks). Replaces: r123::Philox4x32::operator()
 * as called at RandBLAS/random_gen.hh:107,135. */
int rb_philox_words(const uint32_t ctr[4], const uint32_t key[2], int64_t n_blocks, uint32_t* out, void* stream);

/* (g0[i], g1[i]) = boxmuller(w0[i], w1[i]): the float transform the Gaussian family applies to two Philox words
 * (r123::boxmuller as called at RandBLAS/random_gen.hh:62-74). For checks of the device libm emulation on chosen words. */
int rb_boxmuller_words(int64_t n, const uint32_t* w0, const uint32_t* w1, float* g0, float* g1, void* stream);

/* ---- K1: fill_dense ----
 * Replaces RandBLAS::fill_dense_unpacked (RandBLAS/dense_skops.hh:563-606) and through it
 * fill_dense(D, buff, seed) (:623-626), fill_dense(S) (:649-658), dense::fill_dense_submat_impl (:96-170).
 * Writes the n_rows x n_cols window at (ro_s, co_s) of the sample of DenseDist(D_rows, D_cols, family,
 * major_axis) in `layout`, with leading dimension ld (0 = packed, the reference's behaviour). */
int rb_fill_dense_f32(char layout, int64_t D_rows, int64_t D_cols, char family, char major_axis, int64_t n_rows,
                      int64_t n_cols, int64_t ro_s, int64_t co_s, float* buff, int64_t ld, const uint32_t ctr[4],
                      const uint32_t key[2], uint32_t next_ctr[4], void* stream);
int rb_fill_dense_f64(char layout, int64_t D_rows, int64_t D_cols, char family, char major_axis, int64_t n_rows,
                      int64_t n_cols, int64_t ro_s, int64_t co_s, double* buff, int64_t ld, const uint32_t ctr[4],
                      const uint32_t key[2], uint32_t next_ctr[4], void* stream);

/* ---- K3a: SASO index/sign generation ----
 * Replaces RandBLAS::fill_sparse_unpacked_nosub, SASO branch (RandBLAS/sparse_skops.hh:515-533), fill_sparse(S)
 * (:587-603) and sparse::repeated_fisher_yates (:51-106). val_bytes in {4, 8} (float/double), idx_bytes in {4, 8}.
 * vals/rows/cols hold full_nnz = vec_nnz * dim_minor entries. *nnz (host) receives full_nnz. */
int rb_fill_sparse_saso(int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4], const uint32_t key[2],
                        void* vals, int val_bytes, void* rows, void* cols, int idx_bytes, int64_t* nnz,
                        uint32_t next_ctr[4], void* stream);
/* ---- K3c: LASO (Axis::Long) generation ----
 * Replaces the long-axis branch of RandBLAS::fill_sparse_unpacked_nosub (RandBLAS/sparse_skops.hh:534-564) with
 * sample_indices_iid_uniform (RandBLAS/util.hh:515-547) and laso_merge_long_axis_vector_coo_data
 * (sparse_skops.hh:453-491). vals/rows/cols hold full_nnz = vec_nnz * dim_minor entries; *nnz (host) receives the
 * number actually written (repeated indices inside a vector are merged). Entries of a vector come in
 * first-occurrence order (the reference's order for vectors with repeats is std::unordered_map's). Synchronises. */
int rb_fill_sparse_laso(int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4], const uint32_t key[2],
                        void* vals, int val_bytes, void* rows, void* cols, int idx_bytes, int64_t* nnz,
                        uint32_t next_ctr[4], void* stream);
/* public repeated_fisher_yates(k, n, r, samples, state) (RandBLAS/sparse_skops.hh:259-264) */
int rb_repeated_fisher_yates(int64_t k, int64_t n, int64_t r, void* samples, int idx_bytes, const uint32_t ctr[4],
                             const uint32_t key[2], uint32_t next_ctr[4], void* stream);

/* ---- index-sampling utilities ----
 * rb_sample_indices_iid_uniform replaces RandBLAS::sample_indices_iid_uniform<T, sint_t, WriteRademachers>
 * (RandBLAS/util.hh:515-560): k samples from the uniform distribution over {0..n-1} into samples[k] (idx_bytes in
 * {4, 8}); rademachers[k] (val_bytes in {4, 8}) receives the signs, or is NULL for the overload without them
 * (:556-559; it then uses one Philox word per sample instead of two). next_ctr = ctr + ceil(k / (2 or 4)).
 * rb_sample_indices_iid replaces RandBLAS::sample_indices_iid<T, sint_t> (:490-513): k samples from the CDF cdf[n]
 * (val_bytes in {4, 8}) by std::lower_bound semantics; next_ctr = ctr + ceil(k / 4).
 * rb_weights_to_cdf_* replaces RandBLAS::weights_to_cdf<T> (:459-473): w[n] overwritten in place by the normalised
 * running sum of max(w, 0) (summed serially in T, like the reference); RB_ERR_ARG if a weight is below
 * error_if_below or the total is below sqrt(n) * eps (the entries before the failing one have been overwritten, as in
 * the reference). Synchronises. */
int rb_sample_indices_iid_uniform(int64_t n, int64_t k, void* samples, int idx_bytes, void* rademachers, int val_bytes,
                                  const uint32_t ctr[4], const uint32_t key[2], uint32_t next_ctr[4], void* stream);
int rb_sample_indices_iid(int64_t n, const void* cdf, int val_bytes, int64_t k, void* samples, int idx_bytes,
                          const uint32_t ctr[4], const uint32_t key[2], uint32_t next_ctr[4], void* stream);
int rb_weights_to_cdf_f32(int64_t n, float* w, float error_if_below, void* stream);
int rb_weights_to_cdf_f64(int64_t n, double* w, double error_if_below, void* stream);

/* ---- K2: dense operator applied to dense data ----
 * Replaces dense::lskge3 (RandBLAS/skge.hh:154-202) / dense::rskge3 (:307-355), i.e. the DenseSkOp
 * overloads of sketch_general (:799-821, :947-968, :1073-1097, :1175-1199) and sketch_vector (skve.hh:141-164).
 *   left : B(d x n) = alpha * op(S[ro_s:, co_s:])(d x m) * op(A)(m x n) + beta * B
 *   right: B(m x d) = alpha * op(A)(m x n) * op(S[ro_s:, co_s:])(n x d) + beta * B
 * S is the sample of DenseDist(D_rows, D_cols, family, major_axis) at (ctr, key). If S_buff is NULL the
 * operator is regenerated tile by tile inside the kernel and never touches HBM; otherwise S_buff is the
 * filled operator (S.buff, in the operator's natural layout, leading dimension dim_major) and is read. */
int rb_lskge3_f32(char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, float alpha, int64_t D_rows,
                  int64_t D_cols, char family, char major_axis, const uint32_t ctr[4], const uint32_t key[2],
                  const float* S_buff, int64_t ro_s, int64_t co_s, const float* A, int64_t lda, float beta, float* B,
                  int64_t ldb, void* stream);
int rb_lskge3_f64(char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, double alpha, int64_t D_rows,
                  int64_t D_cols, char family, char major_axis, const uint32_t ctr[4], const uint32_t key[2],
                  const double* S_buff, int64_t ro_s, int64_t co_s, const double* A, int64_t lda, double beta,
                  double* B, int64_t ldb, void* stream);
int rb_rskge3_f32(char layout, char opA, char opS, int64_t m, int64_t d, int64_t n, float alpha, const float* A,
                  int64_t lda, int64_t D_rows, int64_t D_cols, char family, char major_axis, const uint32_t ctr[4],
                  const uint32_t key[2], const float* S_buff, int64_t ro_s, int64_t co_s, float beta, float* B,
                  int64_t ldb, void* stream);
int rb_rskge3_f64(char layout, char opA, char opS, int64_t m, int64_t d, int64_t n, double alpha, const double* A,
                  int64_t lda, int64_t D_rows, int64_t D_cols, char family, char major_axis, const uint32_t ctr[4],
                  const uint32_t key[2], const double* S_buff, int64_t ro_s, int64_t co_s, double beta, double* B,
                  int64_t ldb, void* stream);

/* ---- K3b: SASO operator applied to dense data ----
 * Replaces sparse::lskges (RandBLAS/skge.hh:465-492) / sparse::rskges (:598-626) and the left_spmm COO path
 * under them (sparse_data/spmm_dispatch.hh:52-179, coo_spmm_impl.hh:53-105). The operator is
 * SparseDist(D_rows, D_cols, vec_nnz, Short) at (ctr, key); its (row, sign) lists are regenerated in
 * registers, the COO arrays are never materialised. */
int rb_lskges_f32(char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, float alpha, int64_t D_rows,
                  int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s,
                  int64_t co_s, const float* A, int64_t lda, float beta, float* B, int64_t ldb, void* stream);
int rb_lskges_f64(char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, double alpha, int64_t D_rows,
                  int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s,
                  int64_t co_s, const double* A, int64_t lda, double beta, double* B, int64_t ldb, void* stream);
int rb_rskges_f32(char layout, char opA, char opS, int64_t m, int64_t d, int64_t n, float alpha, const float* A,
                  int64_t lda, int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4],
                  const uint32_t key[2], int64_t ro_s, int64_t co_s, float beta, float* B, int64_t ldb, void* stream);
int rb_rskges_f64(char layout, char opA, char opS, int64_t m, int64_t d, int64_t n, double alpha, const double* A,
                  int64_t lda, int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4],
                  const uint32_t key[2], int64_t ro_s, int64_t co_s, double beta, double* B, int64_t ldb, void* stream);
/* Same contraction for an ALREADY SAMPLED sparse operator given as COO arrays (S.vals/rows/cols with S.nnz >= 0,
 * skge.hh:489-490): generic COO x dense. idx_bytes in {4, 8}. */
int rb_coo_apply_f32(int side_left, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, float alpha,
                     int64_t S_rows, int64_t S_cols, int64_t nnz, const float* vals, const void* rows,
                     const void* cols, int idx_bytes, int64_t ro_s, int64_t co_s, const float* A, int64_t lda,
                     float beta, float* B, int64_t ldb, void* stream);
int rb_coo_apply_f64(int side_left, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, double alpha,
                     int64_t S_rows, int64_t S_cols, int64_t nnz, const double* vals, const void* rows,
                     const void* cols, int idx_bytes, int64_t ro_s, int64_t co_s, const double* A, int64_t lda,
                     double beta, double* B, int64_t ldb, void* stream);

/* ---- symmetry check in front of sketch_symmetric ----
 * Replaces util::require_symmetric (RandBLAS/util.hh:128-148), called by sketch_symmetric (RandBLAS/sksy.hh:159-176,
 * 294-312) before it forwards to sketch_general: fails (RB_ERR_ARG, message naming the first offending pair) if
 * |A(i,j) - A(j,i)| > (|A(i,j)| + |A(j,i)| + 1) * tol for some i < j; tol < 0 skips the check. Synchronises. */
int rb_require_symmetric_f32(char layout, const float* A, int64_t n, int64_t lda, float tol, void* stream);
int rb_require_symmetric_f64(char layout, const double* A, int64_t n, int64_t lda, double tol, void* stream);

/* ---- sparse data matrix times dense matrix ----
 * Replaces sparse_data::left_spmm / right_spmm (RandBLAS/sparse_data/spmm_dispatch.hh:52-219; the kernels in
 * coo_spmm_impl.hh, csr_spmm_impl.hh, csc_spmm_impl.hh), the public entry points RandLAPACK calls on sparse data.
 *   side_left = 1: C(d x n) = alpha * op(A_sp[ro_a:, co_a:])(d x m) * op(B)(m x n) + beta * C      (left_spmm)
 *   side_left = 0: C(m x d) = alpha * op(B)(m x n) * op(A_sp[ro_a:, co_a:])(n x d) + beta * C      (right_spmm)
 * fmt: 0 CSR (idx0 = rowptr, idx1 = colidxs), 1 CSC (idx0 = rowidxs, idx1 = colptr), 2 COO (idx0 = rows, idx1 = cols).
 * CSR/CSC take no submatrix (ro_a = co_a = 0, exact dimensions, spmm_dispatch.hh:99-107). index_base must be Zero. */
int rb_spmm_f32(int side_left, int fmt, char layout, char opA, char opB, int64_t d, int64_t n, int64_t m, float alpha,
                int64_t A_rows, int64_t A_cols, int64_t nnz, const float* vals, const void* idx0, const void* idx1,
                int idx_bytes, int64_t ro_a, int64_t co_a, const float* B, int64_t ldb, float beta, float* C, int64_t ldc,
                void* stream);
int rb_spmm_f64(int side_left, int fmt, char layout, char opA, char opB, int64_t d, int64_t n, int64_t m, double alpha,
                int64_t A_rows, int64_t A_cols, int64_t nnz, const double* vals, const void* idx0, const void* idx1,
                int idx_bytes, int64_t ro_a, int64_t co_a, const double* B, int64_t ldb, double beta, double* C,
                int64_t ldc, void* stream);

/* ---- sparse format conversions on the device ----
 * Replaces coo_to_csr / coo_to_csc (RandBLAS/sparse_data/conversions.hh:79-121 with COOMatrix::sort_arrays and
 * sorted_idxs_to_compressed_ptr, sparse_data/base.hh:279-301): to_csc = 0 compresses the rows, 1 the columns.
 * out_vals / out_idx hold nnz entries ordered by (major, minor) index, out_ptr n_major + 1 offsets. */
int rb_coo_to_compressed(int to_csc, int64_t n_rows, int64_t n_cols, int64_t nnz, const void* vals, int val_bytes,
                         const void* rows, const void* cols, int idx_bytes, void* out_vals, void* out_idx, void* out_ptr,
                         void* stream);
/* csr_to_coo / csc_to_coo (conversions.hh:49-75): ptr (n_major + 1 offsets) expanded to one major index per entry;
 * the other two COO arrays are the compressed matrix's own vals and index array. */
int rb_expand_ptr(int64_t n_major, const void* ptr, int64_t nnz, void* out_idx, int idx_bytes, void* stream);

/* ---- K4: dense operator applied to sparse data ----
 * Replaces sparse_data::lsksp3 (RandBLAS/sparse_data/sksp.hh:132-182) / rsksp3 (:277-326), i.e. sketch_sparse
 * (:418-437, :520-539) and the right_spmm/left_spmm kernels under them (spmm_dispatch.hh:52-219,
 * csc_spmm_impl.hh, csr_spmm_impl.hh, coo_spmm_impl.hh).
 *   left : B(d x n) = alpha * op(S[ro_s:, co_s:]) * op(A_sp[ro_a:, co_a:]) + beta * B
 *   right: B(m x d) = alpha * op(A_sp[ro_a:, co_a:]) * op(S[ro_s:, co_s:]) + beta * B
 * fmt 0 = CSR (idx0 = rowptr[A_rows+1], idx1 = colidxs[nnz]); 1 = CSC (idx0 = rowidxs[nnz], idx1 = colptr[A_cols+1]);
 * 2 = COO (idx0 = rows[nnz], idx1 = cols[nnz]). Zero-based indices, idx_bytes in {4, 8}. */
int rb_lsksp3_f32(int fmt, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, float alpha,
                  int64_t D_rows, int64_t D_cols, char family, char major_axis, const uint32_t ctr[4],
                  const uint32_t key[2], int64_t ro_s, int64_t co_s, int64_t A_rows, int64_t A_cols, int64_t nnz,
                  const float* vals, const void* idx0, const void* idx1, int idx_bytes, int64_t ro_a, int64_t co_a,
                  float beta, float* B, int64_t ldb, void* stream);
int rb_lsksp3_f64(int fmt, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, double alpha,
                  int64_t D_rows, int64_t D_cols, char family, char major_axis, const uint32_t ctr[4],
                  const uint32_t key[2], int64_t ro_s, int64_t co_s, int64_t A_rows, int64_t A_cols, int64_t nnz,
                  const double* vals, const void* idx0, const void* idx1, int idx_bytes, int64_t ro_a, int64_t co_a,
                  double beta, double* B, int64_t ldb, void* stream);
int rb_rsksp3_f32(int fmt, char layout, char opA, char opS, int64_t m, int64_t d, int64_t n, float alpha,
                  int64_t A_rows, int64_t A_cols, int64_t nnz, const float* vals, const void* idx0, const void* idx1,
                  int idx_bytes, int64_t ro_a, int64_t co_a, int64_t D_rows, int64_t D_cols, char family,
                  char major_axis, const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s, int64_t co_s,
                  float beta, float* B, int64_t ldb, void* stream);
int rb_rsksp3_f64(int fmt, char layout, char opA, char opS, int64_t m, int64_t d, int64_t n, double alpha,
                  int64_t A_rows, int64_t A_cols, int64_t nnz, const double* vals, const void* idx0, const void* idx1,
                  int idx_bytes, int64_t ro_a, int64_t co_a, int64_t D_rows, int64_t D_cols, char family,
                  char major_axis, const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s, int64_t co_s,
                  double beta, double* B, int64_t ldb, void* stream);

/* ---- random sparse matrices and column partition ----
 * rb_random_coo_* replaces RandBLAS::sparse_data::random_coo<T, sint_t> (RandBLAS/sparse_data/random_matrix.hh:290-355):
 * an m x n COO matrix in CSR sort order whose entries are stored independently with probability `density`
 * (geometric skips over the row-major linearised index, PhiloxStream :64-121) with iid N(0,1) values. The
 * reference's sequential stream is reproduced bit for bit in parallel: stored entries 2b and 2b+1 consume exactly the
 * four words of Philox block ctr + b (skip, Box-Muller pair, skip). Two-call protocol: *nnz always receives the
 * number of stored entries and next_ctr the reference's returned state; vals/rows/cols are written only for
 * entries below `capacity` (call with capacity 0 to size the arrays). *ambiguous (may be NULL) counts skips whose
 * log(1-u)/log(1-p) fell within 8 ulp of an integer, where the device's and the host's double-precision log could
 * round differently: 0 (the overwhelmingly common case) certifies equality with the reference. density in [0, 1).
 * Synchronises. The row-restarting random_csr / random_csc streams (:136-288) are sequential by construction; the
 * host layers build CSR / CSC from this matrix with rb_sorted_idxs_to_ptr (same distribution, not the same stream). */
int rb_random_coo_f32(int64_t m, int64_t n, double density, const uint32_t ctr[4], const uint32_t key[2], int64_t capacity,
                      float* vals, void* rows, void* cols, int idx_bytes, int64_t* nnz, uint32_t next_ctr[4],
                      int64_t* ambiguous, void* stream);
int rb_random_coo_f64(int64_t m, int64_t n, double density, const uint32_t ctr[4], const uint32_t key[2], int64_t capacity,
                      double* vals, void* rows, void* cols, int idx_bytes, int64_t* nnz, uint32_t next_ctr[4],
                      int64_t* ambiguous, void* stream);
/* sorted_idxs_to_compressed_ptr (RandBLAS/sparse_data/base.hh:279-301): ptr[i] = number of entries of the sorted
 * index array idxs[nnz] that are < i, i in [0, n_major]. Device pointers. */
int rb_sorted_idxs_to_ptr(int64_t n_major, int64_t nnz, const void* idxs, int idx_bytes, void* ptr, void* stream);
/* Column block [c0, c1) of a CSR matrix as a new n_rows x (c1 - c0) CSR matrix (column indices shifted by -c0, order
 * inside each row kept): the partition step in front of a column-sharded sketch_sparse. Device
 * pointers. out_rowptr[n_rows + 1] and *nnz_out are always written; out_vals / out_colidxs only if capacity >= the
 * block's entry count (call with capacity 0 to size them). Synchronises. */
int rb_csr_column_block(int64_t n_rows, int64_t n_cols, int64_t nnz, const void* vals, int val_bytes, const void* rowptr,
                        const void* colidxs, int idx_bytes, int64_t c0, int64_t c1, int64_t capacity, void* out_vals,
                        void* out_rowptr, void* out_colidxs, int64_t* nnz_out, void* stream);

/* ---- multi-GPU: the m-sharded left sketch and its one exchange step ----
 * Distributes the reference's own blocked form of a left sketch (RandBLAS/skge.hh:174-181; rtd/source/tutorial/
 * sketch_updates.rst:198-213: row blocks of A against column blocks of S selected with (ro_s, co_s), block products
 * accumulated with beta = 1): the m_total rows of op(A) are split into nranks contiguous blocks whose starts are
 * multiples of 4 (rb_mshard_block), GPU g holds block g as A_local (m_local x n in `layout`, leading dimension lda)
 * and regenerates only the matching columns of the operator; the d x n partial products are summed over NVLink by
 * one NCCL collective inside the library:
 *   mode 0: reduce-scatter -- B_out (d*n / nranks entries, requires d*n % nranks == 0) is this rank's contiguous
 *           slice, in memory order, of the packed result (ldb = d for ColMajor, n for RowMajor);
 *   mode 1: all-reduce     -- B_out (d*n entries) is the whole packed result on every rank.
 * B_out = alpha * op(S[ro_s:, co_s:]) * op(A) + beta * B_out. All data pointers are DEVICE pointers of the
 * communicator's GPU, which must be the current device; the call is ordered on `stream` and asynchronous.
 * The sum order differs from the single-GPU kernel's, so results agree to rounding (1e-12 double / 1e-5 float
 * relative Frobenius), not bit for bit. NCCL is loaded at run time (dlopen libnccl.so.2); a communicator of one
 * rank needs no NCCL.
 *
 * One process per GPU: rank 0 calls rb_comm_unique_id, ships the RB_COMM_ID_BYTES bytes to the other ranks by any
 * means (MPI, a file, torch.distributed), every rank calls rb_comm_init_rank with its GPU current.
 * One process for all GPUs: rb_comm_init(ndev, devices or NULL = 0..ndev-1, comms[ndev]) and the _all entry point,
 * which takes one A_local / lda / B_out / stream per GPU and issues the collective as one NCCL group. */
#define RB_COMM_ID_BYTES 128
typedef struct rb_comm* rb_comm_t;
int rb_comm_unique_id(void* id128);
int rb_comm_init_rank(int nranks, int rank, const void* id128, rb_comm_t* comm);
int rb_comm_init(int ndev, const int* devices, rb_comm_t* comms);
int rb_comm_destroy(rb_comm_t comm);
/* info = {nranks, rank, device, NCCL version code (0 if NCCL is not loaded)} */
int rb_comm_info(rb_comm_t comm, int64_t info[4]);
/* rows [start, start + count) of op(A) owned by `rank` (pure arithmetic) */
int rb_mshard_block(int64_t m_total, int nranks, int rank, int64_t* start, int64_t* count);
int rb_lskge3_mshard_f32(rb_comm_t comm, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m_total,
                         float alpha, int64_t D_rows, int64_t D_cols, char family, char major_axis,
                         const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s, int64_t co_s, const float* A_local,
                         int64_t lda, float beta, float* B_out, int mode, void* stream);
int rb_lskge3_mshard_f64(rb_comm_t comm, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m_total,
                         double alpha, int64_t D_rows, int64_t D_cols, char family, char major_axis,
                         const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s, int64_t co_s, const double* A_local,
                         int64_t lda, double beta, double* B_out, int mode, void* stream);
int rb_lskge3_mshard_all_f32(int ndev, const rb_comm_t* comms, char layout, char opS, char opA, int64_t d, int64_t n,
                             int64_t m_total, float alpha, int64_t D_rows, int64_t D_cols, char family, char major_axis,
                             const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s, int64_t co_s,
                             const float* const* A_local, const int64_t* lda, float beta, float* const* B_out, int mode,
                             void* const* streams);
int rb_lskge3_mshard_all_f64(int ndev, const rb_comm_t* comms, char layout, char opS, char opA, int64_t d, int64_t n,
                             int64_t m_total, double alpha, int64_t D_rows, int64_t D_cols, char family, char major_axis,
                             const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s, int64_t co_s,
                             const double* const* A_local, const int64_t* lda, double beta, double* const* B_out,
                             int mode, void* const* streams);
/* The same for a SASO operator SparseDist(D_rows, D_cols, vec_nnz, Short) (sparse::lskges, RandBLAS/skge.hh:465-492, on
 * row block M_g of op(A) with the operator window shifted by start(M_g) along the contraction index -- SURVEY.md 8(e),
 * "SASO apply, m-sharded": the d x n partials (2 MB at BASELINE.json configs[3]) meet in the same reduce-scatter /
 * all-reduce. Minor-axis vectors of the operator are independent (counter i * vec_nnz + j, sparse_skops.hh:72-102), so
 * any block start would do; rb_mshard_block is used for both operator kinds. */
int rb_lskges_mshard_f32(rb_comm_t comm, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m_total,
                         float alpha, int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4],
                         const uint32_t key[2], int64_t ro_s, int64_t co_s, const float* A_local, int64_t lda, float beta,
                         float* B_out, int mode, void* stream);
int rb_lskges_mshard_f64(rb_comm_t comm, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m_total,
                         double alpha, int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4],
                         const uint32_t key[2], int64_t ro_s, int64_t co_s, const double* A_local, int64_t lda,
                         double beta, double* B_out, int mode, void* stream);
int rb_lskges_mshard_all_f32(int ndev, const rb_comm_t* comms, char layout, char opS, char opA, int64_t d, int64_t n,
                             int64_t m_total, float alpha, int64_t D_rows, int64_t D_cols, int64_t vec_nnz,
                             const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s, int64_t co_s,
                             const float* const* A_local, const int64_t* lda, float beta, float* const* B_out, int mode,
                             void* const* streams);
int rb_lskges_mshard_all_f64(int ndev, const rb_comm_t* comms, char layout, char opS, char opA, int64_t d, int64_t n,
                             int64_t m_total, double alpha, int64_t D_rows, int64_t D_cols, int64_t vec_nnz,
                             const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s, int64_t co_s,
                             const double* const* A_local, const int64_t* lda, double beta, double* const* B_out,
                             int mode, void* const* streams);

/* ---- tuning / introspection (not part of the reference's surface) ----
 * rb_set_option("dense_path", v): 0 = auto (tensor-core kernels where the shape allows), 1 = force the
 * generic SIMT kernel. Other switches select among kernels that compute the same result (kept for measurement):
 * "tc_cluster" (0 / 1 / 2: 2-CTA cluster mode of the float tensor-core kernel: never / where it pays / whenever
 * possible), "tc_splits", "tc_halves", "saso_path" (0 auto / 1 atomic kernel / 2 binned kernel), "saso_fill_path" (0 auto: thread per vector for
 * vec_nnz = 2, 4, 8, 16 / 1 warp per vector / 2 lane per entry), "saso_bin_path" (binning pass of the apply: 0 auto / 1 lane per entry),
 * "fill_rep" (Gaussian fill_dense of long vectors: 1 = tiles of four passes of 8 Philox blocks per thread, the default / 0 = one pass),
 * "fill_unroll" (Uniform float fill_dense of long vectors: 1 = 16 Philox blocks per thread and tile, the default / 0 = 4),
 * "saso_rows" (apply kernel: 1 = a lane owns a whole row of the register tile, the default / 0 = an 8-lane group owns 8 rows),
 * "spdata_path" (1 = the deterministic column-owner kernel), "h2d_chunk_mb" (block size of the host-pointer sketch
 * pipeline; the first two blocks are a quarter and a half of it), "tc_pair" (CTA pairs, cta_group::2: 0 never / 1 where it pays / 2 whenever possible), "tc_materialise"
 * (Gaussian float operators generated once per K panel: 0 never / 1 with >= 2 column tiles / 2 always),
 * "tc_ymn" (float kernel, Q-contiguous data: 0 = tiles fed to the tensor core as an MN-major operand, the default; 1 = tiles
 * transposed by the generator warps), "tc_xmn" (materialised operators contiguous along the rows of op(S), i.e. filled
 * Axis::Short operators: 1 = tensor-core / DMMA kernels, the default; 0 = generic kernel),
 * "dmma_materialise" (the same for double: 1 = Gaussian operators with K >= 4096, the default; 0 = fused kernel),
 * "dmma_panel_mb" (size of that panel, default 2048).
 * rb_get_counter("kernel_launches" | "tensor_core_launches" | "saso_owner_launches") counts kernels this library
 * launched. */
int rb_set_option(const char* name, int64_t value);
/* current value of an option (0 for an unknown name) */
int64_t rb_get_option(const char* name);
int64_t rb_get_counter(const char* name);

#ifdef __cplusplus
}
#endif
#endif /* RANDBLAS_B200_H */
