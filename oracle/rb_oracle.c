/* TEST INFRASTRUCTURE -- not product code. See rb_oracle.h for scope and pinning status.
 *
 * Plain-C restatement of the reference's sketching hot path. Each function cites the reference
 * file:line it follows (paths relative to the reference root). Arithmetic that decides bits
 * (Philox, counter adds, uneg11/u01, Box-Muller call sequence, Fisher-Yates) is restated
 * operation for operation; the GEMM-like products are restated as the mathematical contract
 * (B = alpha*op(S)*op(A) + beta*B) with double accumulation, because the reference delegates
 * them to a vendor BLAS whose summation order is unspecified.
 *
 * Compiled with -ffp-contract=off so that no multiply-add here is fused behind our back
 * (every product that feeds an add in the bit-exact parts is by a power of two anyway).
 */
#define _GNU_SOURCE
#include "rb_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10. Random123 philox.h (un-vendored dependency, unpinned HEAD in the reference CI:
 * .github/workflows/core-linux.yaml:34-39); call sites RandBLAS/random_gen.hh:107,135,
 * RandBLAS/sparse_skops.hh:77. Pinned by test/test_basic_rng/r123_kat_vectors.txt:19-21.
 * ------------------------------------------------------------------------------------------ */
void rbo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t pa = (uint64_t) 0xD2511F53u * c0;
        uint64_t pb = (uint64_t) 0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(pb >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t) pb;
        uint32_t n2 = (uint32_t)(pa >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t) pa;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Threefry4x32-20: only used to reproduce the uneg11/u01 histogram pins of
 * test/test_basic_rng/test_r123.cc:568-671. Pinned by r123_kat_vectors.txt:50-52. */
static uint32_t rotl32(uint32_t x, unsigned n) { return (x << n) | (x >> (32 - n)); }
void rbo_threefry4x32_20(const uint32_t ctr[4], const uint32_t key[4], uint32_t out[4]) {
    static const unsigned ra[8] = {10, 11, 13, 23, 6, 17, 25, 18};
    static const unsigned rb[8] = {26, 21, 27, 5, 20, 11, 10, 20};
    uint32_t ks[5];
    ks[4] = 0x1BD11BDAu;
    for (int i = 0; i < 4; ++i) { ks[i] = key[i]; ks[4] ^= key[i]; }
    uint32_t x0 = ctr[0] + ks[0], x1 = ctr[1] + ks[1], x2 = ctr[2] + ks[2], x3 = ctr[3] + ks[3];
    for (unsigned r = 0; r < 20; ++r) {
        if ((r & 1u) == 0) {
            x0 += x1; x1 = rotl32(x1, ra[r % 8]); x1 ^= x0;
            x2 += x3; x3 = rotl32(x3, rb[r % 8]); x3 ^= x2;
        } else {
            x0 += x3; x3 = rotl32(x3, ra[r % 8]); x3 ^= x0;
            x2 += x1; x1 = rotl32(x1, rb[r % 8]); x1 ^= x2;
        }
        if (r % 4 == 3) {
            unsigned s = r / 4 + 1;
            x0 += ks[s % 5]; x1 += ks[(s + 1) % 5]; x2 += ks[(s + 2) % 5]; x3 += ks[(s + 3) % 5] + s;
        }
    }
    out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
}

/* 128-bit little-endian counter += 64-bit n. Random123 array.h `incr`; semantics pinned by
 * test/test_basic_rng/test_r123.cc:710-797. */
void rbo_ctr_incr(uint32_t ctr[4], uint64_t n) {
    uint64_t lo = ((uint64_t) ctr[1] << 32) | ctr[0];
    uint64_t hi = ((uint64_t) ctr[3] << 32) | ctr[2];
    uint64_t s = lo + n;
    if (s < lo) hi += 1;
    ctr[0] = (uint32_t) s; ctr[1] = (uint32_t)(s >> 32);
    ctr[2] = (uint32_t) hi; ctr[3] = (uint32_t)(hi >> 32);
}

/* RNGState(uint64 k): counter = 0, key = 0 then key.incr(k). RandBLAS/base.hh:116-119;
 * pinned by test_r123.cc:679-698. */
void rbo_rngstate_from_u64(uint64_t k, uint32_t ctr[4], uint32_t key[2]) {
    ctr[0] = ctr[1] = ctr[2] = ctr[3] = 0;
    key[0] = (uint32_t) k;
    key[1] = (uint32_t)(k >> 32);
}

/* Random123 uniform.hpp, float instantiations (RandBLAS/random_gen.hh:127-136 picks float for
 * 32-bit counters regardless of the matrix scalar type). factor = 1/(float(INT32_MAX)+1) = 2^-31. */
float rbo_uneg11_f32(uint32_t w) {
    const float factor = 0x1p-31f, halffactor = 0x1p-32f;
    return (float)(int32_t) w * factor + halffactor;
}
float rbo_u01_f32(uint32_t w) {
    const float factor = 0x1p-32f, halffactor = 0x1p-33f;
    return (float) w * factor + halffactor;
}
/* Random123 boxmuller.hpp (float, host branch); called pairwise by boxmulall,
 * RandBLAS/random_gen.hh:62-74. */
void rbo_boxmuller_f32(uint32_t u0, uint32_t u1, float* x, float* y) {
    const float PIf = 3.1415926535897932f;
    float s, c;
    sincosf(PIf * rbo_uneg11_f32(u0), &s, &c);
    float r = sqrtf(-2.f * logf(rbo_u01_f32(u1)));
    *x = s * r;
    *y = c * r;
}

/* ------------------------------------------------------------------------------------------
 * DenseDist. RandBLAS/dense_skops.hh:187-199 (natural_layout), :300-328 (constructor).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t n_rows, n_cols, dim_major, dim_minor;
    char family, axis, natural_layout;
} ddist_t;

static int mk_ddist(int64_t n_rows, int64_t n_cols, char family, char axis, ddist_t* D) {
    if (n_rows <= 0 || n_cols <= 0) return 1;
    if (family != 'G' && family != 'U') return 1;
    if (axis != 'S' && axis != 'L') return 1;
    int64_t mx = n_rows > n_cols ? n_rows : n_cols, mn = n_rows < n_cols ? n_rows : n_cols;
    D->n_rows = n_rows; D->n_cols = n_cols; D->family = family; D->axis = axis;
    D->dim_major = (axis == 'L') ? mx : mn;
    D->dim_minor = (axis == 'L') ? mn : mx;
    int is_wide = n_rows < n_cols, fa_long = axis == 'L';
    D->natural_layout = (is_wide && fa_long) ? 'R' : (is_wide ? 'C' : (fa_long ? 'C' : 'R'));
    return 0;
}

int rbo_dense_dist_info(int64_t n_rows, int64_t n_cols, char family, char axis, int64_t info[3], double* iso) {
    ddist_t D;
    if (mk_ddist(n_rows, n_cols, family, axis, &D)) return 1;
    info[0] = D.dim_major; info[1] = D.dim_minor; info[2] = D.natural_layout;
    *iso = pow((double) D.dim_minor, -0.5);
    return 0;
}

/* dense::compute_next_state, RandBLAS/dense_skops.hh:172-185 */
int rbo_dense_next_state(int64_t n_rows, int64_t n_cols, char family, char axis, const uint32_t ctr[4],
                         uint32_t next_ctr[4]) {
    ddist_t D;
    if (mk_ddist(n_rows, n_cols, family, axis, &D)) return 1;
    int64_t blocks_per_major_vec = (D.dim_major + 3) / 4;
    memcpy(next_ctr, ctr, 16);
    rbo_ctr_incr(next_ctr, (uint64_t)(blocks_per_major_vec * D.dim_minor));
    return 0;
}

/* One Philox block -> 4 float samples of the given family. r123ext::uneg11 / boxmul,
 * RandBLAS/random_gen.hh:85-137. */
static void gen4(char family, const uint32_t ctr[4], const uint32_t key[2], float out[4]) {
    uint32_t w[4];
    rbo_philox4x32_10(ctr, key, w);
    if (family == 'U') {
        for (int i = 0; i < 4; ++i) out[i] = rbo_uneg11_f32(w[i]);
    } else {
        rbo_boxmuller_f32(w[0], w[1], &out[0], &out[1]);
        rbo_boxmuller_f32(w[2], w[3], &out[2], &out[3]);
    }
}

/* ------------------------------------------------------------------------------------------
 * SparseDist. RandBLAS/sparse_skops.hh:108-114, :205-224.
 * ------------------------------------------------------------------------------------------ */
int rbo_sparse_dist_info(int64_t n_rows, int64_t n_cols, int64_t vec_nnz, char axis, int64_t info[3], double* iso) {
    if (n_rows <= 0 || n_cols <= 0 || vec_nnz <= 0) return 1;
    int64_t mx = n_rows > n_cols ? n_rows : n_cols, mn = n_rows < n_cols ? n_rows : n_cols;
    int64_t dim_major = (axis == 'S') ? mn : mx;
    int64_t dim_minor = n_rows + n_cols - dim_major;
    if (vec_nnz > dim_major) return 1;
    info[0] = dim_major; info[1] = dim_minor; info[2] = vec_nnz * dim_minor;
    if (axis == 'S') *iso = pow((double) vec_nnz, -0.5);
    else *iso = sqrt(((double) dim_major) / (vec_nnz * ((double) dim_minor)));
    return 0;
}

/* compute_next_state(SparseDist), RandBLAS/sparse_skops.hh:266-283 */
int rbo_sparse_next_state(int64_t n_rows, int64_t n_cols, int64_t vec_nnz, char axis, const uint32_t ctr[4],
                          uint32_t next_ctr[4]) {
    int64_t info[3]; double iso;
    if (rbo_sparse_dist_info(n_rows, n_cols, vec_nnz, axis, info, &iso)) return 1;
    int64_t mx = n_rows > n_cols ? n_rows : n_cols, mn = n_rows < n_cols ? n_rows : n_cols;
    int64_t num_mavec, incrs;
    if (axis == 'S') { num_mavec = mx; incrs = vec_nnz; }
    else { num_mavec = mn; incrs = (int64_t) ceil((double) vec_nnz / 2.0); }
    memcpy(next_ctr, ctr, 16);
    rbo_ctr_incr(next_ctr, (uint64_t)(num_mavec * incrs));
    return 0;
}

static void store_idx(void* arr, int idx_bytes, int64_t pos, int64_t v) {
    if (idx_bytes == 4) ((int32_t*) arr)[pos] = (int32_t) v; else ((int64_t*) arr)[pos] = v;
}
static void store_val(void* arr, int val_bytes, int64_t pos, double v) {
    if (val_bytes == 4) ((float*) arr)[pos] = (float) v; else ((double*) arr)[pos] = v;
}

/* sparse::repeated_fisher_yates, RandBLAS/sparse_skops.hh:51-106. For each of dim_minor vectors:
 * vec_nnz Fisher-Yates steps over a work permutation of [0, dim_major), one Philox block per
 * step at counter seed + i*vec_nnz + j; ell = j + w[0] % (dim_major - j); sign from w[1] parity;
 * the permutation is put back to identity after each vector (:92-102). */
static int fisher_yates(const uint32_t ctr[4], const uint32_t key[2], int64_t vec_nnz, int64_t dim_major,
                        int64_t dim_minor, void* idxs_major, void* idxs_minor, int idx_bytes, void* vals,
                        int val_bytes, uint32_t next_ctr[4]) {
    if (vec_nnz > dim_major) return 1;
    int64_t* work = (int64_t*) malloc(sizeof(int64_t) * (size_t) dim_major);
    int64_t* piv = (int64_t*) malloc(sizeof(int64_t) * (size_t) vec_nnz);
    int64_t* drawn = (int64_t*) malloc(sizeof(int64_t) * (size_t) vec_nnz);
    for (int64_t j = 0; j < dim_major; ++j) work[j] = j;
    for (int64_t i = 0; i < dim_minor; ++i) {
        int64_t offset = i * vec_nnz;
        uint32_t c[4];
        memcpy(c, ctr, 16);
        rbo_ctr_incr(c, (uint64_t) offset);
        for (int64_t j = 0; j < vec_nnz; ++j) {
            uint32_t w[4];
            rbo_philox4x32_10(c, key, w);
            int64_t ell = j + (int64_t)(w[0] % (uint64_t)(dim_major - j));
            piv[j] = ell;
            int64_t picked = work[ell];
            work[ell] = work[j];
            work[j] = picked;
            drawn[j] = picked;
            store_idx(idxs_major, idx_bytes, offset + j, picked);
            if (vals) store_val(vals, val_bytes, offset + j, (w[1] % 2 == 0) ? 1.0 : -1.0);
            if (idxs_minor) store_idx(idxs_minor, idx_bytes, offset + j, i);
            rbo_ctr_incr(c, 1);
        }
        for (int64_t j = vec_nnz - 1; j >= 0; --j) { /* undo the swaps, last first */
            int64_t ell = piv[j];
            work[j] = work[ell];
            work[ell] = drawn[j];
        }
    }
    if (next_ctr) { memcpy(next_ctr, ctr, 16); rbo_ctr_incr(next_ctr, (uint64_t)(dim_minor * vec_nnz)); }
    free(work); free(piv); free(drawn);
    return 0;
}

/* public repeated_fisher_yates(k, n, r, samples, state), RandBLAS/sparse_skops.hh:259-264 */
int rbo_repeated_fisher_yates(int64_t k, int64_t n, int64_t r, void* samples, int idx_bytes, const uint32_t ctr[4],
                              const uint32_t key[2], uint32_t next_ctr[4]) {
    return fisher_yates(ctr, key, k, n, r, samples, NULL, idx_bytes, NULL, 0, next_ctr);
}

/* ---- index-sampling utilities, RandBLAS/util.hh:459-560 (serial restatement of the reference's loops) ---- */

/* sample_indices_iid_uniform<T, sint_t, WriteRademachers>, util.hh:515-547: one stream of Philox blocks from `ctr`;
 * len_c = 4 words per block, rounded down to a multiple of two when Rademachers are written (:521-525). */
int rbo_sample_indices_iid_uniform(int64_t n, int64_t k, void* samples, int idx_bytes, void* rademachers, int val_bytes,
                                   const uint32_t ctr[4], const uint32_t key[2], uint32_t next_ctr[4]) {
    uint32_t c[4], w[4];
    memcpy(c, ctr, 16);
    rbo_philox4x32_10(c, key, w);
    const int len_c = 4;                       /* 2 * (4 / 2) with Rademachers: also 4 */
    int rv_index = 0;
    const double dN = (double) n;
    for (int64_t i = 0; i < k; ++i) {
        const double u01 = ((double) rbo_uneg11_f32(w[rv_index]) + 1.0) / 2.0;      /* uneg11_to_u01<double>, :475-478 */
        if (idx_bytes == 4) ((int32_t*) samples)[i] = (int32_t) ((double) (int32_t) dN * u01);   /* :532-533 */
        else ((int64_t*) samples)[i] = (int64_t) ((double) (int64_t) dN * u01);
        rv_index += 1;
        if (rademachers) {
            const double r = (rbo_uneg11_f32(w[rv_index]) >= 0) ? 1.0 : -1.0;       /* :536 */
            store_val(rademachers, val_bytes, i, r);
            rv_index += 1;
        }
        if (rv_index == len_c) {
            rbo_ctr_incr(c, 1);
            rbo_philox4x32_10(c, key, w);
            rv_index = 0;
        }
    }
    if (0 < rv_index) rbo_ctr_incr(c, 1);
    if (next_ctr) memcpy(next_ctr, c, 16);
    return 0;
}

/* sample_indices_iid<T, sint_t>, util.hh:490-513: u = ((T) x + 1) / 2 in T, std::lower_bound on the CDF */
int rbo_sample_indices_iid(int64_t n, const void* cdf, int val_bytes, int64_t k, void* samples, int idx_bytes,
                           const uint32_t ctr[4], const uint32_t key[2], uint32_t next_ctr[4]) {
    uint32_t c[4], w[4];
    memcpy(c, ctr, 16);
    rbo_philox4x32_10(c, key, w);
    int rv_index = 0;
    for (int64_t i = 0; i < k; ++i) {
        int64_t lo = 0, len = n;
        if (val_bytes == 4) {
            const float u = (rbo_uneg11_f32(w[rv_index]) + 1.0f) / 2.0f;
            const float* f = (const float*) cdf;
            while (len > 0) { int64_t h = len / 2; if (f[lo + h] < u) { lo += h + 1; len -= h + 1; } else len = h; }
        } else {
            const double u = ((double) rbo_uneg11_f32(w[rv_index]) + 1.0) / 2.0;
            const double* f = (const double*) cdf;
            while (len > 0) { int64_t h = len / 2; if (f[lo + h] < u) { lo += h + 1; len -= h + 1; } else len = h; }
        }
        store_idx(samples, idx_bytes, i, lo);
        rv_index += 1;
        if (rv_index == 4) {
            rbo_ctr_incr(c, 1);
            rbo_philox4x32_10(c, key, w);
            rv_index = 0;
        }
    }
    if (0 < rv_index) rbo_ctr_incr(c, 1);
    if (next_ctr) memcpy(next_ctr, c, 16);
    return 0;
}

/* weights_to_cdf<T>, util.hh:459-473. Returns 1 where the reference throws (w keeps what had been written). */
int rbo_weights_to_cdf_f32(int64_t n, float* w, float error_if_below) {
    float sum = 0.0f;
    for (int64_t i = 0; i < n; ++i) {
        float val = w[i];
        if (!(val >= error_if_below)) return 1;
        val = (val < 0.0f) ? 0.0f : val;
        sum += val;
        w[i] = sum;
    }
    if (!(sum >= ((float) sqrt((double) n)) * FLT_EPSILON)) return 1;
    const float a = 1.0f / sum;
    for (int64_t i = 0; i < n; ++i) w[i] = w[i] * a;        /* blas::scal */
    return 0;
}
int rbo_weights_to_cdf_f64(int64_t n, double* w, double error_if_below) {
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double val = w[i];
        if (!(val >= error_if_below)) return 1;
        val = (val < 0.0) ? 0.0 : val;
        sum += val;
        w[i] = sum;
    }
    if (!(sum >= ((double) sqrt((double) n)) * DBL_EPSILON)) return 1;
    const double a = 1.0 / sum;
    for (int64_t i = 0; i < n; ++i) w[i] = w[i] * a;
    return 0;
}

/* fill_sparse_unpacked_nosub, SASO branch, RandBLAS/sparse_skops.hh:515-533: the short-axis index
 * array is `rows` when n_rows <= n_cols, else `cols`. */
int rbo_fill_sparse_saso(int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4], const uint32_t key[2],
                         void* vals, int val_bytes, void* rows, void* cols, int idx_bytes, int64_t* nnz,
                         uint32_t next_ctr[4]) {
    int64_t info[3]; double iso;
    if (rbo_sparse_dist_info(D_rows, D_cols, vec_nnz, 'S', info, &iso)) return 1;
    void* idxs_short = (D_rows <= D_cols) ? rows : cols;
    void* idxs_long = (D_rows <= D_cols) ? cols : rows;
    int rc = fisher_yates(ctr, key, vec_nnz, info[0], info[1], idxs_short, idxs_long, idx_bytes, vals, val_bytes,
                          next_ctr);
    if (rc == 0) *nnz = vec_nnz * info[1];
    return rc;
}

/* fill_sparse_unpacked_nosub, LASO branch (Axis::Long), RandBLAS/sparse_skops.hh:534-564, with
 * sample_indices_iid_uniform<T, sint_t, WriteRademachers = true> (RandBLAS/util.hh:515-547) and
 * laso_merge_long_axis_vector_coo_data (sparse_skops.hh:453-491).
 * The state runs sequentially through the vectors: per vector, pairs of uneg11 floats are consumed from
 * consecutive Philox blocks (len_c = 4 -> two (index, sign) pairs per block); a partially used block is skipped
 * at the end of the vector (util.hh:545), so vector i starts at seed + i * ceil(vec_nnz / 2).
 * Repeated indices: one entry, value sqrt(count) * (sign of the first occurrence), in T arithmetic.
 * ORDER: the reference rewrites a vector that has repeats in std::unordered_map iteration order (a property of
 * the C++ library, sparse_skops.hh:484-488); this restatement uses first-occurrence order, callers compare such
 * vectors as sets. Vectors without repeats keep draw order in both. */
int rbo_fill_sparse_laso(int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4], const uint32_t key[2],
                         void* vals, int val_bytes, void* rows, void* cols, int idx_bytes, int64_t* nnz,
                         uint32_t next_ctr[4]) {
    int64_t info[3]; double iso;
    if (rbo_sparse_dist_info(D_rows, D_cols, vec_nnz, 'L', info, &iso)) return 1;
    const int64_t dim_major = info[0], dim_minor = info[1];
    void* idxs_short = (D_rows <= D_cols) ? rows : cols;
    void* idxs_long = (D_rows <= D_cols) ? cols : rows;
    int64_t* idx = (int64_t*) malloc(sizeof(int64_t) * (size_t) vec_nnz);
    int* neg = (int*) malloc(sizeof(int) * (size_t) vec_nnz);
    uint32_t c[4];
    memcpy(c, ctr, 16);
    const double dN = (double) dim_major;
    int64_t total = 0;
    for (int64_t i = 0; i < dim_minor; ++i) {
        uint32_t w[4];
        rbo_philox4x32_10(c, key, w);
        int rv_index = 0;
        for (int64_t j = 0; j < vec_nnz; ++j) {                       /* util.hh:530-544 */
            const double u01 = ((double) rbo_uneg11_f32(w[rv_index]) + 1.0) / 2.0;
            idx[j] = (int64_t) ((double) (int64_t) dN * u01);
            rv_index += 1;
            neg[j] = (rbo_uneg11_f32(w[rv_index]) >= 0) ? 0 : 1;
            rv_index += 1;
            if (rv_index == 4) { rbo_ctr_incr(c, 1); rbo_philox4x32_10(c, key, w); rv_index = 0; }
        }
        if (rv_index > 0) rbo_ctr_incr(c, 1);                         /* util.hh:545 */
        for (int64_t j = 0; j < vec_nnz; ++j) {                       /* merge, first-occurrence order */
            int first = 1;
            for (int64_t t = 0; t < j; ++t) if (idx[t] == idx[j]) { first = 0; break; }
            if (!first) continue;
            int64_t count = 1;
            for (int64_t t = j + 1; t < vec_nnz; ++t) if (idx[t] == idx[j]) ++count;
            store_idx(idxs_long, idx_bytes, total, idx[j]);
            store_idx(idxs_short, idx_bytes, total, i);
            if (val_bytes == 4) ((float*) vals)[total] = sqrtf((float) count) * (neg[j] ? -1.0f : 1.0f);
            else ((double*) vals)[total] = sqrt((double) count) * (neg[j] ? -1.0 : 1.0);
            ++total;
        }
    }
    *nnz = total;
    if (next_ctr) memcpy(next_ctr, c, 16);
    free(idx); free(neg);
    return 0;
}

int rbo_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#endif
    (void) n;
    return 0;
}
int rbo_get_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* typed part, instantiated for float and double */
#define T float
#define SFX(name) name##_f32
#include "rb_oracle_typed.inc"
#undef T
#undef SFX
#define T double
#define SFX(name) name##_f64
#include "rb_oracle_typed.inc"
#undef T
#undef SFX
