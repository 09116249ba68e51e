// TEST INFRASTRUCTURE -- not product code. Nothing under randblas_b200/ may link or load this.
//
// extern "C" face of the UNMODIFIED reference (header-only RandBLAS, read where it lies under
// /root/reference) so that Python tests / golden generation / the CPU-baseline leg of bench.py
// can run the reference's own loops: fill_dense_submat_impl, repeated_fisher_yates, lskge3,
// lskges, lsksp3, left_spmm ... Only the two un-vendored leaf dependencies (Random123, BLAS++)
// come from oracle/shims/. Built by oracle/Makefile into oracle/_ref/librb_ref.so.
//
// Every function returns 0 on success, 1 if the reference threw RandBLAS::Error (argument
// validation, exceptions.hh:57-95), 2 for any other exception; message via rbref_last_error().

#include <RandBLAS.hh>
#include <RandBLAS/sparse_data/random_matrix.hh>
#include <omp.h>
#include <cstdint>
#include <cstring>
#include <string>

using RandBLAS::RNGState;
using RandBLAS::DenseDist;
using RandBLAS::DenseSkOp;
using RandBLAS::SparseDist;
using RandBLAS::SparseSkOp;
using RandBLAS::ScalarDist;
using RandBLAS::Axis;
using blas::Layout;
using blas::Op;
using RNG = r123::Philox4x32;

static thread_local std::string g_err;

#define RBREF_TRY try {
#define RBREF_CATCH                                              \
    } catch (const RandBLAS::Error& e) { g_err = e.what(); return 1; } \
      catch (const std::exception& e) { g_err = e.what(); return 2; }  \
      catch (...) { g_err = "unknown exception"; return 2; }           \
    return 0;

static RNGState<RNG> mk_state(const uint32_t* ctr, const uint32_t* key) {
    RNGState<RNG> s;
    for (int i = 0; i < 4; ++i) s.counter.v[i] = ctr[i];
    for (int i = 0; i < 2; ++i) s.key.v[i] = key[i];
    return s;
}
static void put_ctr(const RNGState<RNG>& s, uint32_t* ctr_out) {
    if (ctr_out) for (int i = 0; i < 4; ++i) ctr_out[i] = s.counter.v[i];
}
static DenseDist mk_dd(int64_t r, int64_t c, char family, char axis) {
    return DenseDist(r, c, (ScalarDist) family, (Axis) axis);
}
static SparseDist mk_sd(int64_t r, int64_t c, int64_t vec_nnz, char axis) {
    return SparseDist(r, c, vec_nnz, (Axis) axis);
}

extern "C" {

const char* rbref_last_error(void) { return g_err.c_str(); }
const char* rbref_blas_config(void) { return scipy_openblas_get_config(); }
int rbref_set_threads(int n) { omp_set_num_threads(n); scipy_openblas_set_num_threads(n); return 0; }
int rbref_get_threads(void) { return omp_get_max_threads(); }
int rbref_get_blas_threads(void) { return scipy_openblas_get_num_threads(); }

int rbref_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    RBREF_TRY
    auto s = mk_state(ctr, key);
    RNG rng;
    auto r = rng(s.counter, s.key);
    for (int i = 0; i < 4; ++i) out[i] = r.v[i];
    RBREF_CATCH
}

int rbref_ctr_incr(uint32_t* ctr, uint64_t n) {
    RBREF_TRY
    RNG::ctr_type c;
    for (int i = 0; i < 4; ++i) c.v[i] = ctr[i];
    c.incr(n);
    for (int i = 0; i < 4; ++i) ctr[i] = c.v[i];
    RBREF_CATCH
}

int rbref_rngstate_from_u64(uint64_t k, uint32_t* ctr, uint32_t* key) {
    RBREF_TRY
    RNGState<RNG> s(k);
    for (int i = 0; i < 4; ++i) ctr[i] = s.counter.v[i];
    for (int i = 0; i < 2; ++i) key[i] = s.key.v[i];
    RBREF_CATCH
}

int rbref_uneg11_f32(const uint32_t* ctr, const uint32_t* key, float* out) {
    RBREF_TRY
    auto s = mk_state(ctr, key);
    RNG rng;
    auto r = r123ext::uneg11::generate(rng, s.counter, s.key);
    for (int i = 0; i < 4; ++i) out[i] = r[i];
    RBREF_CATCH
}

int rbref_boxmul_f32(const uint32_t* ctr, const uint32_t* key, float* out) {
    RBREF_TRY
    auto s = mk_state(ctr, key);
    RNG rng;
    auto r = r123ext::boxmul::generate(rng, s.counter, s.key);
    for (int i = 0; i < 4; ++i) out[i] = r[i];
    RBREF_CATCH
}

// info[0]=dim_major info[1]=dim_minor info[2]=natural_layout ('R'/'C')
int rbref_dense_dist_info(int64_t n_rows, int64_t n_cols, char family, char axis, int64_t* info, double* iso) {
    RBREF_TRY
    auto D = mk_dd(n_rows, n_cols, family, axis);
    info[0] = D.dim_major; info[1] = D.dim_minor; info[2] = (int64_t)(char) D.natural_layout;
    *iso = D.isometry_scale;
    RBREF_CATCH
}

int rbref_dense_next_state(int64_t n_rows, int64_t n_cols, char family, char axis, const uint32_t* ctr,
                           const uint32_t* key, uint32_t* next_ctr) {
    RBREF_TRY
    auto D = mk_dd(n_rows, n_cols, family, axis);
    DenseSkOp<float, RNG> S(D, mk_state(ctr, key));
    put_ctr(S.next_state, next_ctr);
    RBREF_CATCH
}

// info[0]=dim_major info[1]=dim_minor info[2]=full_nnz
int rbref_sparse_dist_info(int64_t n_rows, int64_t n_cols, int64_t vec_nnz, char axis, int64_t* info, double* iso) {
    RBREF_TRY
    auto D = mk_sd(n_rows, n_cols, vec_nnz, axis);
    info[0] = D.dim_major; info[1] = D.dim_minor; info[2] = D.full_nnz;
    *iso = D.isometry_scale;
    RBREF_CATCH
}

int rbref_sparse_next_state(int64_t n_rows, int64_t n_cols, int64_t vec_nnz, char axis, const uint32_t* ctr,
                            const uint32_t* key, uint32_t* next_ctr) {
    RBREF_TRY
    auto D = mk_sd(n_rows, n_cols, vec_nnz, axis);
    SparseSkOp<float, RNG> S(D, mk_state(ctr, key));
    put_ctr(S.next_state, next_ctr);
    RBREF_CATCH
}

} // extern "C"

// ---------------------------------------------------------------------------------------------
template <typename T>
static int fill_dense_unpacked_t(char layout, int64_t Dr, int64_t Dc, char family, char axis, int64_t n_rows,
                                 int64_t n_cols, int64_t ro_s, int64_t co_s, T* buff, const uint32_t* ctr,
                                 const uint32_t* key, uint32_t* next_ctr) {
    RBREF_TRY
    auto D = mk_dd(Dr, Dc, family, axis);
    auto next = RandBLAS::fill_dense_unpacked((Layout) layout, D, n_rows, n_cols, ro_s, co_s, buff, mk_state(ctr, key));
    put_ctr(next, next_ctr);
    RBREF_CATCH
}

template <typename T, typename sint_t>
static int fill_sparse_t(int64_t Dr, int64_t Dc, int64_t vec_nnz, char axis, const uint32_t* ctr, const uint32_t* key,
                         T* vals, sint_t* rows, sint_t* cols, int64_t* nnz, uint32_t* next_ctr) {
    RBREF_TRY
    auto D = mk_sd(Dr, Dc, vec_nnz, axis);
    auto next = RandBLAS::fill_sparse_unpacked_nosub(D, *nnz, vals, rows, cols, mk_state(ctr, key));
    put_ctr(next, next_ctr);
    RBREF_CATCH
}

template <typename sint_t>
static int rfy_t(int64_t k, int64_t n, int64_t r, sint_t* samples, const uint32_t* ctr, const uint32_t* key,
                 uint32_t* next_ctr) {
    RBREF_TRY
    auto next = RandBLAS::repeated_fisher_yates(k, n, r, samples, mk_state(ctr, key));
    put_ctr(next, next_ctr);
    RBREF_CATCH
}

// index-sampling utilities, RandBLAS/util.hh:459-560
template <typename T, typename sint_t>
static int iid_uniform_t(int64_t n, int64_t k, sint_t* samples, T* rad, const uint32_t* ctr, const uint32_t* key,
                         uint32_t* next_ctr) {
    RBREF_TRY
    auto st = mk_state(ctr, key);
    auto next = rad ? RandBLAS::sample_indices_iid_uniform<T, sint_t, true>(n, k, samples, rad, st)
                    : RandBLAS::sample_indices_iid_uniform(n, k, samples, st);
    put_ctr(next, next_ctr);
    RBREF_CATCH
}
template <typename T, typename sint_t>
static int iid_t(int64_t n, const T* cdf, int64_t k, sint_t* samples, const uint32_t* ctr, const uint32_t* key,
                 uint32_t* next_ctr) {
    RBREF_TRY
    auto next = RandBLAS::sample_indices_iid(n, cdf, k, samples, mk_state(ctr, key));
    put_ctr(next, next_ctr);
    RBREF_CATCH
}
template <typename T>
static int w2cdf_t(int64_t n, T* w, T error_if_below) {
    RBREF_TRY
    RandBLAS::weights_to_cdf(n, w, error_if_below);
    RBREF_CATCH
}

// dense operator, left/right sketch. prefill != 0 => fill_dense(S) first (exercises the buff != nullptr path)
template <typename T>
static int skge_dense_t(int side_left, char layout, char op1, char op2, int64_t d, int64_t n, int64_t m, T alpha,
                        int64_t Dr, int64_t Dc, char family, char axis, const uint32_t* ctr, const uint32_t* key,
                        int prefill, int64_t ro_s, int64_t co_s, const T* A, int64_t lda, T beta, T* B, int64_t ldb) {
    RBREF_TRY
    DenseSkOp<T, RNG> S(mk_dd(Dr, Dc, family, axis), mk_state(ctr, key));
    if (prefill) RandBLAS::fill_dense(S);
    if (side_left) {
        // (layout, opS, opA, d, n, m, alpha, S, ro_s, co_s, A, lda, beta, B, ldb)   skge.hh:799-821
        RandBLAS::sketch_general((Layout) layout, (Op) op1, (Op) op2, d, n, m, alpha, S, ro_s, co_s, A, lda, beta, B, ldb);
    } else {
        // (layout, opA, opS, m, d, n, alpha, A, lda, S, ro_s, co_s, beta, B, ldb)   skge.hh:947-968
        // here the caller passes (d, n, m) in the *reference's right-sketch order* (m, d, n)
        RandBLAS::sketch_general((Layout) layout, (Op) op1, (Op) op2, d, n, m, alpha, A, lda, S, ro_s, co_s, beta, B, ldb);
    }
    RBREF_CATCH
}

template <typename T>
static int skge_sparse_t(int side_left, char layout, char op1, char op2, int64_t d, int64_t n, int64_t m, T alpha,
                         int64_t Dr, int64_t Dc, int64_t vec_nnz, char axis, const uint32_t* ctr, const uint32_t* key,
                         int prefill, int64_t ro_s, int64_t co_s, const T* A, int64_t lda, T beta, T* B, int64_t ldb) {
    RBREF_TRY
    SparseSkOp<T, RNG> S(mk_sd(Dr, Dc, vec_nnz, axis), mk_state(ctr, key));
    if (prefill) RandBLAS::fill_sparse(S);
    if (side_left) {
        RandBLAS::sketch_general((Layout) layout, (Op) op1, (Op) op2, d, n, m, alpha, S, ro_s, co_s, A, lda, beta, B, ldb);
    } else {
        RandBLAS::sketch_general((Layout) layout, (Op) op1, (Op) op2, d, n, m, alpha, A, lda, S, ro_s, co_s, beta, B, ldb);
    }
    RBREF_CATCH
}

// sparse data times dense operator. fmt: 0 = CSR (idx0=rowptr, idx1=colidxs), 1 = CSC (idx0=rowidxs, idx1=colptr),
// 2 = COO (idx0=rows, idx1=cols). int64 indices (the reference's default sint_t).
template <typename T>
static int sksp_t(int side_left, int fmt, char layout, char op1, char op2, int64_t d, int64_t n, int64_t m, T alpha,
                  int64_t Dr, int64_t Dc, char family, char axis, const uint32_t* ctr, const uint32_t* key, int prefill,
                  int64_t ro_s, int64_t co_s, int64_t Ar, int64_t Ac, int64_t nnz, T* vals, int64_t* idx0, int64_t* idx1,
                  T beta, T* B, int64_t ldb) {
    RBREF_TRY
    using namespace RandBLAS::sparse_data;
    DenseSkOp<T, RNG> S(mk_dd(Dr, Dc, family, axis), mk_state(ctr, key));
    if (prefill) RandBLAS::fill_dense(S);
    auto run = [&](auto& Asp) {
        if (side_left)
            RandBLAS::sketch_sparse((Layout) layout, (Op) op1, (Op) op2, d, n, m, alpha, S, ro_s, co_s, Asp, beta, B, ldb);
        else
            RandBLAS::sketch_sparse((Layout) layout, (Op) op1, (Op) op2, d, n, m, alpha, Asp, S, ro_s, co_s, beta, B, ldb);
    };
    if (fmt == 0) { CSRMatrix<T, int64_t> Asp(Ar, Ac, nnz, vals, idx0, idx1); run(Asp); }
    else if (fmt == 1) { CSCMatrix<T, int64_t> Asp(Ar, Ac, nnz, vals, idx0, idx1); run(Asp); }
    else { COOMatrix<T, int64_t> Asp(Ar, Ac, nnz, vals, idx0, idx1); run(Asp); }
    RBREF_CATCH
}

template <typename T>
static int skve_dense_t(char opS, int64_t d, int64_t m, T alpha, int64_t Dr, int64_t Dc, char family, char axis,
                        const uint32_t* ctr, const uint32_t* key, int64_t ro_s, int64_t co_s, const T* x, int64_t incx,
                        T beta, T* y, int64_t incy) {
    RBREF_TRY
    DenseSkOp<T, RNG> S(mk_dd(Dr, Dc, family, axis), mk_state(ctr, key));
    RandBLAS::sketch_vector((Op) opS, d, m, alpha, S, ro_s, co_s, x, incx, beta, y, incy);
    RBREF_CATCH
}

template <typename T>
static int skve_sparse_t(char opS, int64_t d, int64_t m, T alpha, int64_t Dr, int64_t Dc, int64_t vec_nnz, char axis,
                         const uint32_t* ctr, const uint32_t* key, int64_t ro_s, int64_t co_s, const T* x, int64_t incx,
                         T beta, T* y, int64_t incy) {
    RBREF_TRY
    SparseSkOp<T, RNG> S(mk_sd(Dr, Dc, vec_nnz, axis), mk_state(ctr, key));
    RandBLAS::sketch_vector((Op) opS, d, m, alpha, S, ro_s, co_s, x, incx, beta, y, incy);
    RBREF_CATCH
}

extern "C" {

#define DEF_FOR_T(T, sfx)                                                                                              \
    int rbref_fill_dense_unpacked_##sfx(char layout, int64_t Dr, int64_t Dc, char family, char axis, int64_t n_rows,   \
                                        int64_t n_cols, int64_t ro_s, int64_t co_s, T* buff, const uint32_t* ctr,      \
                                        const uint32_t* key, uint32_t* next_ctr) {                                     \
        return fill_dense_unpacked_t<T>(layout, Dr, Dc, family, axis, n_rows, n_cols, ro_s, co_s, buff, ctr, key,      \
                                        next_ctr);                                                                     \
    }                                                                                                                  \
    int rbref_fill_sparse_##sfx##_i32(int64_t Dr, int64_t Dc, int64_t vec_nnz, char axis, const uint32_t* ctr,         \
                                      const uint32_t* key, T* vals, int32_t* rows, int32_t* cols, int64_t* nnz,        \
                                      uint32_t* next_ctr) {                                                            \
        return fill_sparse_t<T, int32_t>(Dr, Dc, vec_nnz, axis, ctr, key, vals, rows, cols, nnz, next_ctr);            \
    }                                                                                                                  \
    int rbref_fill_sparse_##sfx##_i64(int64_t Dr, int64_t Dc, int64_t vec_nnz, char axis, const uint32_t* ctr,         \
                                      const uint32_t* key, T* vals, int64_t* rows, int64_t* cols, int64_t* nnz,        \
                                      uint32_t* next_ctr) {                                                            \
        return fill_sparse_t<T, int64_t>(Dr, Dc, vec_nnz, axis, ctr, key, vals, rows, cols, nnz, next_ctr);            \
    }                                                                                                                  \
    int rbref_sketch_general_dense_##sfx(int side_left, char layout, char op1, char op2, int64_t d, int64_t n,         \
                                         int64_t m, T alpha, int64_t Dr, int64_t Dc, char family, char axis,           \
                                         const uint32_t* ctr, const uint32_t* key, int prefill, int64_t ro_s,          \
                                         int64_t co_s, const T* A, int64_t lda, T beta, T* B, int64_t ldb) {           \
        return skge_dense_t<T>(side_left, layout, op1, op2, d, n, m, alpha, Dr, Dc, family, axis, ctr, key, prefill,   \
                               ro_s, co_s, A, lda, beta, B, ldb);                                                      \
    }                                                                                                                  \
    int rbref_sketch_general_sparse_##sfx(int side_left, char layout, char op1, char op2, int64_t d, int64_t n,        \
                                          int64_t m, T alpha, int64_t Dr, int64_t Dc, int64_t vec_nnz, char axis,      \
                                          const uint32_t* ctr, const uint32_t* key, int prefill, int64_t ro_s,         \
                                          int64_t co_s, const T* A, int64_t lda, T beta, T* B, int64_t ldb) {          \
        return skge_sparse_t<T>(side_left, layout, op1, op2, d, n, m, alpha, Dr, Dc, vec_nnz, axis, ctr, key, prefill, \
                                ro_s, co_s, A, lda, beta, B, ldb);                                                     \
    }                                                                                                                  \
    int rbref_sketch_sparse_##sfx(int side_left, int fmt, char layout, char op1, char op2, int64_t d, int64_t n,       \
                                  int64_t m, T alpha, int64_t Dr, int64_t Dc, char family, char axis,                  \
                                  const uint32_t* ctr, const uint32_t* key, int prefill, int64_t ro_s, int64_t co_s,   \
                                  int64_t Ar, int64_t Ac, int64_t nnz, T* vals, int64_t* idx0, int64_t* idx1, T beta,  \
                                  T* B, int64_t ldb) {                                                                 \
        return sksp_t<T>(side_left, fmt, layout, op1, op2, d, n, m, alpha, Dr, Dc, family, axis, ctr, key, prefill,    \
                         ro_s, co_s, Ar, Ac, nnz, vals, idx0, idx1, beta, B, ldb);                                     \
    }                                                                                                                  \
    int rbref_sketch_vector_dense_##sfx(char opS, int64_t d, int64_t m, T alpha, int64_t Dr, int64_t Dc, char family,  \
                                        char axis, const uint32_t* ctr, const uint32_t* key, int64_t ro_s,             \
                                        int64_t co_s, const T* x, int64_t incx, T beta, T* y, int64_t incy) {          \
        return skve_dense_t<T>(opS, d, m, alpha, Dr, Dc, family, axis, ctr, key, ro_s, co_s, x, incx, beta, y, incy);  \
    }                                                                                                                  \
    int rbref_sketch_vector_sparse_##sfx(char opS, int64_t d, int64_t m, T alpha, int64_t Dr, int64_t Dc,              \
                                         int64_t vec_nnz, char axis, const uint32_t* ctr, const uint32_t* key,         \
                                         int64_t ro_s, int64_t co_s, const T* x, int64_t incx, T beta, T* y,           \
                                         int64_t incy) {                                                               \
        return skve_sparse_t<T>(opS, d, m, alpha, Dr, Dc, vec_nnz, axis, ctr, key, ro_s, co_s, x, incx, beta, y,       \
                                incy);                                                                                 \
    }

DEF_FOR_T(float, f32)
DEF_FOR_T(double, f64)

int rbref_sample_indices_iid_uniform_f32_i32(int64_t n, int64_t k, int32_t* s, float* rad, const uint32_t* ctr, const uint32_t* key, uint32_t* nx) { return iid_uniform_t<float, int32_t>(n, k, s, rad, ctr, key, nx); }
int rbref_sample_indices_iid_uniform_f32_i64(int64_t n, int64_t k, int64_t* s, float* rad, const uint32_t* ctr, const uint32_t* key, uint32_t* nx) { return iid_uniform_t<float, int64_t>(n, k, s, rad, ctr, key, nx); }
int rbref_sample_indices_iid_uniform_f64_i32(int64_t n, int64_t k, int32_t* s, double* rad, const uint32_t* ctr, const uint32_t* key, uint32_t* nx) { return iid_uniform_t<double, int32_t>(n, k, s, rad, ctr, key, nx); }
int rbref_sample_indices_iid_uniform_f64_i64(int64_t n, int64_t k, int64_t* s, double* rad, const uint32_t* ctr, const uint32_t* key, uint32_t* nx) { return iid_uniform_t<double, int64_t>(n, k, s, rad, ctr, key, nx); }
int rbref_sample_indices_iid_f32_i32(int64_t n, const float* cdf, int64_t k, int32_t* s, const uint32_t* ctr, const uint32_t* key, uint32_t* nx) { return iid_t<float, int32_t>(n, cdf, k, s, ctr, key, nx); }
int rbref_sample_indices_iid_f32_i64(int64_t n, const float* cdf, int64_t k, int64_t* s, const uint32_t* ctr, const uint32_t* key, uint32_t* nx) { return iid_t<float, int64_t>(n, cdf, k, s, ctr, key, nx); }
int rbref_sample_indices_iid_f64_i32(int64_t n, const double* cdf, int64_t k, int32_t* s, const uint32_t* ctr, const uint32_t* key, uint32_t* nx) { return iid_t<double, int32_t>(n, cdf, k, s, ctr, key, nx); }
int rbref_sample_indices_iid_f64_i64(int64_t n, const double* cdf, int64_t k, int64_t* s, const uint32_t* ctr, const uint32_t* key, uint32_t* nx) { return iid_t<double, int64_t>(n, cdf, k, s, ctr, key, nx); }
int rbref_weights_to_cdf_f32(int64_t n, float* w, float eib) { return w2cdf_t<float>(n, w, eib); }
int rbref_weights_to_cdf_f64(int64_t n, double* w, double eib) { return w2cdf_t<double>(n, w, eib); }

int rbref_repeated_fisher_yates_i32(int64_t k, int64_t n, int64_t r, int32_t* samples, const uint32_t* ctr,
                                    const uint32_t* key, uint32_t* next_ctr) {
    return rfy_t<int32_t>(k, n, r, samples, ctr, key, next_ctr);
}
int rbref_repeated_fisher_yates_i64(int64_t k, int64_t n, int64_t r, int64_t* samples, const uint32_t* ctr,
                                    const uint32_t* key, uint32_t* next_ctr) {
    return rfy_t<int64_t>(k, n, r, samples, ctr, key, next_ctr);
}

} // extern "C"

// public left_spmm / right_spmm on sparse DATA (sparse_data/spmm_dispatch.hh:52-219). fmt as in sksp_t. int64 indices.
template <typename T>
static int spmm_t(int side_left, int fmt, char layout, char op1, char op2, int64_t x, int64_t y, int64_t z, T alpha,
                  int64_t Ar, int64_t Ac, int64_t nnz, T* vals, int64_t* idx0, int64_t* idx1, int64_t ro_a, int64_t co_a,
                  const T* B, int64_t ldb, T beta, T* C, int64_t ldc) {
    RBREF_TRY
    using namespace RandBLAS::sparse_data;
    auto run = [&](auto& Asp) {
        if (side_left)   // (layout, opA, opB, d, n, m, alpha, A, ro_a, co_a, B, ldb, beta, C, ldc)
            left_spmm((Layout) layout, (Op) op1, (Op) op2, x, y, z, alpha, Asp, ro_a, co_a, B, ldb, beta, C, ldc);
        else             // (layout, opA(dense), opB(sparse), m, d, n, alpha, A, lda, B, i_off, j_off, beta, C, ldc)
            right_spmm((Layout) layout, (Op) op1, (Op) op2, x, y, z, alpha, B, ldb, Asp, ro_a, co_a, beta, C, ldc);
    };
    if (fmt == 0) { CSRMatrix<T, int64_t> Asp(Ar, Ac, nnz, vals, idx0, idx1); run(Asp); }
    else if (fmt == 1) { CSCMatrix<T, int64_t> Asp(Ar, Ac, nnz, vals, idx0, idx1); run(Asp); }
    else { COOMatrix<T, int64_t> Asp(Ar, Ac, nnz, vals, idx0, idx1); run(Asp); }
    RBREF_CATCH
}

// coo_to_csr / coo_to_csc (sparse_data/conversions.hh:79-121). Outputs: ovals[nnz], oidx[nnz], optr[n_major + 1].
template <typename T>
static int coo_to_compressed_t(int to_csc, int64_t Ar, int64_t Ac, int64_t nnz, T* vals, int64_t* rows, int64_t* cols,
                               T* ovals, int64_t* oidx, int64_t* optr) {
    RBREF_TRY
    using namespace RandBLAS::sparse_data;
    COOMatrix<T, int64_t> coo(Ar, Ac, nnz, vals, rows, cols);
    const int64_t n_major = to_csc ? Ac : Ar;
    for (int64_t i = 0; i <= n_major; ++i) optr[i] = 0;
    if (to_csc) {
        CSCMatrix<T, int64_t> out(Ar, Ac);
        coo_to_csc(coo, out);
        if (out.nnz > 0) {
            std::copy(out.vals, out.vals + out.nnz, ovals);
            std::copy(out.rowidxs, out.rowidxs + out.nnz, oidx);
            std::copy(out.colptr, out.colptr + Ac + 1, optr);
        }
    } else {
        CSRMatrix<T, int64_t> out(Ar, Ac);
        coo_to_csr(coo, out);
        if (out.nnz > 0) {
            std::copy(out.vals, out.vals + out.nnz, ovals);
            std::copy(out.colidxs, out.colidxs + out.nnz, oidx);
            std::copy(out.rowptr, out.rowptr + Ar + 1, optr);
        }
    }
    RBREF_CATCH
}

// csr_to_coo / csc_to_coo (conversions.hh:49-75): only the expanded index array is new
template <typename T>
static int compressed_to_coo_t(int from_csc, int64_t Ar, int64_t Ac, int64_t nnz, T* vals, int64_t* idx, int64_t* ptr,
                               int64_t* expanded) {
    RBREF_TRY
    using namespace RandBLAS::sparse_data;
    COOMatrix<T, int64_t> coo(Ar, Ac);
    if (from_csc) { CSCMatrix<T, int64_t> in(Ar, Ac, nnz, vals, idx, ptr); csc_to_coo(in, coo); std::copy(coo.cols, coo.cols + nnz, expanded); }
    else { CSRMatrix<T, int64_t> in(Ar, Ac, nnz, vals, ptr, idx); csr_to_coo(in, coo); std::copy(coo.rows, coo.rows + nnz, expanded); }
    RBREF_CATCH
}

// random_csr / random_csc / random_coo (sparse_data/random_matrix.hh:136-355). which: 0 CSR, 1 CSC, 2 COO.
// Two-call protocol: capacity < nnz => only *nnz_out and next_ctr are written.
// Outputs follow the C-ABI's (vals, idx0, idx1) convention of sksp_t.
template <typename T>
static int random_sparse_t(int which, int64_t m, int64_t n, double density, const uint32_t* ctr, const uint32_t* key,
                           int64_t capacity, T* vals, int64_t* idx0, int64_t* idx1, int64_t* nnz_out, uint32_t* next_ctr) {
    RBREF_TRY
    using namespace RandBLAS::sparse_data;
    auto st = mk_state(ctr, key);
    if (which == 0) {
        auto [A, nx] = random_csr<T, int64_t>(m, n, density, st);
        put_ctr(nx, next_ctr); *nnz_out = A.nnz;
        if (capacity >= A.nnz) {
            if (A.rowptr) std::copy(A.rowptr, A.rowptr + m + 1, idx0); else for (int64_t i = 0; i <= m; ++i) idx0[i] = 0;
            std::copy(A.vals, A.vals + A.nnz, vals); std::copy(A.colidxs, A.colidxs + A.nnz, idx1);
        }
    } else if (which == 1) {
        auto [A, nx] = random_csc<T, int64_t>(m, n, density, st);
        put_ctr(nx, next_ctr); *nnz_out = A.nnz;
        if (capacity >= A.nnz) {
            if (A.colptr) std::copy(A.colptr, A.colptr + n + 1, idx1); else for (int64_t i = 0; i <= n; ++i) idx1[i] = 0;
            std::copy(A.vals, A.vals + A.nnz, vals); std::copy(A.rowidxs, A.rowidxs + A.nnz, idx0);
        }
    } else {
        auto [A, nx] = random_coo<T, int64_t>(m, n, density, st);
        put_ctr(nx, next_ctr); *nnz_out = A.nnz;
        if (capacity >= A.nnz) {
            std::copy(A.vals, A.vals + A.nnz, vals); std::copy(A.rows, A.rows + A.nnz, idx0); std::copy(A.cols, A.cols + A.nnz, idx1);
        }
    }
    RBREF_CATCH
}

extern "C" {
#define DEF_SP_FOR_T(T, sfx)                                                                                            \
    int rbref_spmm_##sfx(int side_left, int fmt, char layout, char op1, char op2, int64_t x, int64_t y, int64_t z,      \
                         T alpha, int64_t Ar, int64_t Ac, int64_t nnz, T* vals, int64_t* idx0, int64_t* idx1,           \
                         int64_t ro_a, int64_t co_a, const T* B, int64_t ldb, T beta, T* C, int64_t ldc) {              \
        return spmm_t<T>(side_left, fmt, layout, op1, op2, x, y, z, alpha, Ar, Ac, nnz, vals, idx0, idx1, ro_a, co_a,   \
                         B, ldb, beta, C, ldc);                                                                         \
    }                                                                                                                   \
    int rbref_coo_to_compressed_##sfx(int to_csc, int64_t Ar, int64_t Ac, int64_t nnz, T* vals, int64_t* rows,          \
                                      int64_t* cols, T* ovals, int64_t* oidx, int64_t* optr) {                          \
        return coo_to_compressed_t<T>(to_csc, Ar, Ac, nnz, vals, rows, cols, ovals, oidx, optr);                        \
    }                                                                                                                   \
    int rbref_compressed_to_coo_##sfx(int from_csc, int64_t Ar, int64_t Ac, int64_t nnz, T* vals, int64_t* idx,         \
                                      int64_t* ptr, int64_t* expanded) {                                                \
        return compressed_to_coo_t<T>(from_csc, Ar, Ac, nnz, vals, idx, ptr, expanded);                                 \
    }                                                                                                                   \
    int rbref_random_sparse_##sfx(int which, int64_t m, int64_t n, double density, const uint32_t* ctr,                 \
                                  const uint32_t* key, int64_t capacity, T* vals, int64_t* idx0, int64_t* idx1,         \
                                  int64_t* nnz_out, uint32_t* next_ctr) {                                               \
        return random_sparse_t<T>(which, m, n, density, ctr, key, capacity, vals, idx0, idx1, nnz_out, next_ctr);       \
    }
DEF_SP_FOR_T(float, f32)
DEF_SP_FOR_T(double, f64)
}  // extern "C"
