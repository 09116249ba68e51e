// Random sparse test matrices on the device and the column partition of compressed formats
// (SURVEY.md section 8f, ranks 3 and 4).
//
// random_coo (RandBLAS/sparse_data/random_matrix.hh:290-355) walks ONE sequential PhiloxStream (:64-121): a geometric
// skip (one 32-bit word through u01<double> and the host's double-precision log, :113-116) to the next stored
// position of the row-major linearised matrix, then one Gaussian value; Box-Muller yields two values per pair of
// words and the second is cached (:93-106). Because every skip is followed by exactly one value, the word
// consumption is periodic: stored entries 2b and 2b+1 use exactly the four words of Philox block b --
//   word 0: skip of entry 2b, words 1,2: the Box-Muller pair (entry 2b gets the sine value, entry 2b+1 the cached
//   cosine value), word 3: skip of entry 2b+1.
// So entry q is a pure function of (seed counter + q/2, key) and the positions are a prefix sum of the skips:
// the stream is reproduced BIT FOR BIT in parallel (skips -> cub scan -> scatter), with no sequential walk.
// The only host-libm dependence is log(1 - u) in double: the device's log() is within 1 ulp like glibc's, so
// floor(log(1-u) / log(1-p)) can differ only when the quotient lies within a few ulp of an integer; such draws are
// counted and reported (`ambiguous`, expected 1e-12 of the draws), 0 means the output provably equals the reference's.
//
// random_csr / random_csc (:136-288) restart the column walk in every row, which makes the word position of a row a
// function of the number of entries stored before it: inherently sequential. The device versions are defined as the
// CSR / CSC form of random_coo's matrix (same distribution: iid Bernoulli(density) pattern, N(0,1) values), built
// with sorted_idxs_to_compressed_ptr (sparse_data/base.hh:279-301) as a binary search per row.
#include <cub/cub.cuh>
#include <cmath>
#include "../../include/randblas_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace rb {

namespace {

struct RcooState { long long carry; long long nnz; unsigned long long ambiguous; };

// inc[i] = distance from the previous stored position to the one of draw q0 + i (+ the running position for i == 0)
__global__ void __launch_bounds__(256) rcoo_skips_kernel(Ctr128 ctr, PhiloxKey key, long long q0, int nq, double log_1_minus_p,
                                                         long long* __restrict__ inc, RcooState* __restrict__ stt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const long long q = q0 + i;
    const uint4 w = philox4x32_10(ctr_add(ctr, (uint64_t) (q >> 1)), key);
    const uint32_t word = (q & 1) ? w.w : w.x;
    const double u = __fma_rn((double) word, 0x1p-32, 0x1p-33);          // r123::u01<double>(uint32_t)
    const double quo = log(1.0 - u) / log_1_minus_p;                      // PhiloxStream::geometric, :113-116
    const double fl = floor(quo);
    const double gap = fmin(quo - fl, fl + 1.0 - quo);
    if (gap <= fabs(quo) * 0x1p-49) atomicAdd(&stt->ambiguous, 1ull);
    long long v = (long long) fl + (q > 0 ? 1 : 0);
    if (i == 0) v += stt->carry;
    inc[i] = v;
}

template <typename T, typename IDX>
__global__ void __launch_bounds__(256) rcoo_write_kernel(Ctr128 ctr, PhiloxKey key, long long q0, int nq, long long total,
                                                         long long n_cols, long long capacity, const long long* __restrict__ pos,
                                                         T* __restrict__ vals, IDX* __restrict__ rows, IDX* __restrict__ cols,
                                                         const double2* __restrict__ logtab, RcooState* __restrict__ stt,
                                                         long long prev_last) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const long long q = q0 + i, p = pos[i];
    if (p >= total) {
        const long long before = (i == 0) ? prev_last : pos[i - 1];
        if (before < total) stt->nnz = q;                 // the first draw past the end: q entries are stored
        return;
    }
    if (q >= capacity) return;
    const uint4 w = philox4x32_10(ctr_add(ctr, (uint64_t) (q >> 1)), key);
    float g0, g1;
    boxmuller(w.y, w.z, logtab, g0, g1);
    vals[q] = (T) ((q & 1) ? g1 : g0);
    rows[q] = (IDX) (p / n_cols);
    cols[q] = (IDX) (p % n_cols);
}

__global__ void rcoo_carry_kernel(const long long* __restrict__ pos, int nq, RcooState* __restrict__ stt) {
    stt->carry = pos[nq - 1];
}

// ptr[i] = number of entries whose (sorted) major index is < i, i in [0, n_major]
template <typename IDX>
__global__ void __launch_bounds__(256) sorted_to_ptr_kernel(int64_t n_major, int64_t nnz, const IDX* __restrict__ idx,
                                                            IDX* __restrict__ ptr) {
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i <= n_major; i += (int64_t) gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = nnz;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((int64_t) idx[mid] < i) lo = mid + 1; else hi = mid;
        }
        ptr[i] = (IDX) lo;
    }
}

// ---- column block [c0, c1) of a CSR matrix: count, scan, write (one warp per row, order inside a row kept) ----
template <typename IDX>
__global__ void __launch_bounds__(256) colblock_count_kernel(int64_t n_rows, const IDX* __restrict__ rowptr,
                                                             const IDX* __restrict__ colidxs, int64_t c0, int64_t c1,
                                                             unsigned long long* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += nwarps) {
        const int64_t b = (int64_t) rowptr[r], e = (int64_t) rowptr[r + 1];
        int cnt = 0;
        for (int64_t j = b + lane; j < e; j += 32) {
            const int64_t c = (int64_t) colidxs[j];
            cnt += (c >= c0 && c < c1) ? 1 : 0;
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) counts[r] = (unsigned long long) cnt;
    }
    if (warp == 0 && lane == 0) counts[n_rows] = 0;
}

template <typename T, typename IDX>
__global__ void __launch_bounds__(256) colblock_write_kernel(int64_t n_rows, const IDX* __restrict__ rowptr,
                                                             const IDX* __restrict__ colidxs, const T* __restrict__ vals,
                                                             int64_t c0, int64_t c1, const unsigned long long* __restrict__ optr,
                                                             IDX* __restrict__ out_rowptr, IDX* __restrict__ out_cols,
                                                             T* __restrict__ out_vals, int write_entries) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += nwarps) {
        const int64_t b = (int64_t) rowptr[r], e = (int64_t) rowptr[r + 1];
        int64_t o = (int64_t) optr[r];
        if (lane == 0) out_rowptr[r] = (IDX) o;
        if (!write_entries) continue;
        for (int64_t j0 = b; j0 < e; j0 += 32) {
            const int64_t j = j0 + lane;
            int64_t c = -1;
            T v = (T) 0;
            if (j < e) { c = (int64_t) colidxs[j]; v = vals[j]; }
            const bool keep = c >= c0 && c < c1;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int64_t dst = o + __popc(m & ((1u << lane) - 1u));
                out_cols[dst] = (IDX) (c - c0);
                out_vals[dst] = v;
            }
            o += __popc(m);
        }
    }
    if (warp == 0 && lane == 0) out_rowptr[n_rows] = (IDX) optr[n_rows];
}

template <typename T, typename IDX>
int random_coo_t(int64_t m, int64_t n, double density, Ctr128 ctr, PhiloxKey key, int64_t capacity, T* vals, IDX* rows,
                 IDX* cols, int64_t* nnz_out, uint32_t* next_ctr, int64_t* ambiguous_out, cudaStream_t st) {
    const long long total = (long long) m * (long long) n;
    const double log_1_minus_p = std::log(1.0 - density);       // host libm, as the reference (:328)
    const int BATCH = 1 << 22;
    long long* inc = (long long*) workspace(0, (size_t) BATCH * 8, st);
    long long* pos = (long long*) workspace(1, (size_t) BATCH * 8, st);
    RcooState* stt = (RcooState*) workspace(2, sizeof(RcooState), st);
    size_t tmp_bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, inc, pos, BATCH, st);
    void* tmp = workspace(3, tmp_bytes, st);
    const double2* logtab = logf_table_device();
    if (!inc || !pos || !stt || !tmp || !logtab) return fail_cuda(cudaErrorMemoryAllocation, "random_coo workspace");
    RcooState h{0, -1, 0};
    RB_CUDA(cudaMemcpyAsync(stt, &h, sizeof h, cudaMemcpyHostToDevice, st));
    long long q0 = 0, prev_last = -1;
    // expected number of stored entries total * density: size the batches so that small matrices take one pass
    while (true) {
        const double expect_left = ((double) total - (double) (prev_last + 1)) * density;
        long long want = (long long) (expect_left * 1.05 + 6.0 * std::sqrt(expect_left + 1.0) + 64.0);
        const int nq = (int) (want > BATCH ? BATCH : (want < 256 ? 256 : want));
        const unsigned grid = (unsigned) ((nq + 255) / 256);
        rcoo_skips_kernel<<<grid, 256, 0, st>>>(ctr, key, q0, nq, log_1_minus_p, inc, stt);
        RB_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, inc, pos, nq, st));
        rcoo_write_kernel<T, IDX><<<grid, 256, 0, st>>>(ctr, key, q0, nq, total, (long long) n, (long long) capacity, pos,
                                                        vals, rows, cols, logtab, stt, prev_last);
        rcoo_carry_kernel<<<1, 1, 0, st>>>(pos, nq, stt);
        count_launch(4);
        RB_CUDA(cudaGetLastError());
        RB_CUDA(cudaMemcpyAsync(&h, stt, sizeof h, cudaMemcpyDeviceToHost, st));
        RB_CUDA(cudaStreamSynchronize(st));
        prev_last = h.carry;
        q0 += nq;
        if (h.nnz >= 0) break;
    }
    *nnz_out = (int64_t) h.nnz;
    if (ambiguous_out) *ambiguous_out = (int64_t) h.ambiguous;
    // the stream stops right after the skip that passed the end: draw number nnz, in Philox block nnz / 2 (:343-347)
    store_ctr(ctr_add(ctr, (uint64_t) (h.nnz / 2 + 1)), next_ctr);
    return 0;
}

}  // namespace

static bool is_host_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return !(a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged);
}

template <typename T>
static int random_coo_impl(int64_t m, int64_t n, double density, const uint32_t* ctr, const uint32_t* key, int64_t capacity,
                           T* vals, void* rows, void* cols, int idx_bytes, int64_t* nnz, uint32_t* next_ctr,
                           int64_t* ambiguous, void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    if (capacity > 0 && vals && rows && cols && (is_host_ptr(vals) || is_host_ptr(rows) || is_host_ptr(cols))) {
        // host destination (a caller of the CPU reference): generate into device scratch, copy the stored entries back
        RB_REQUIRE(is_host_ptr(vals) && is_host_ptr(rows) && is_host_ptr(cols));
        T* dv = nullptr; void* dr = nullptr; void* dc = nullptr;
        RB_CUDA(cudaMallocAsync(&dv, (size_t) capacity * sizeof(T), st));
        if (cudaMallocAsync(&dr, (size_t) capacity * idx_bytes, st) != cudaSuccess ||
            cudaMallocAsync(&dc, (size_t) capacity * idx_bytes, st) != cudaSuccess) {
            cudaGetLastError(); cudaFreeAsync(dv, st); if (dr) cudaFreeAsync(dr, st);
            return fail_cuda(cudaErrorMemoryAllocation, "random_coo staging");
        }
        int rc = random_coo_impl<T>(m, n, density, ctr, key, capacity, dv, dr, dc, idx_bytes, nnz, next_ctr, ambiguous, stream);
        if (!rc && *nnz > 0) {
            const size_t k = (size_t) (*nnz < capacity ? *nnz : capacity);
            cudaMemcpyAsync(vals, dv, k * sizeof(T), cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(rows, dr, k * idx_bytes, cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(cols, dc, k * idx_bytes, cudaMemcpyDeviceToHost, st);
        }
        cudaFreeAsync(dv, st); cudaFreeAsync(dr, st); cudaFreeAsync(dc, st);
        if (cudaStreamSynchronize(st) != cudaSuccess && !rc) rc = fail_cuda(cudaGetLastError(), "random_coo copy back");
        return rc;
    }
    RB_REQUIRE(density >= 0.0 && density <= 1.0);            // random_matrix.hh:297
    RB_REQUIRE(m >= 0 && n >= 0 && capacity >= 0);
    RB_REQUIRE(ctr != nullptr && key != nullptr && nnz != nullptr);
    RB_REQUIRE(idx_bytes == 4 || idx_bytes == 8);
    RB_REQUIRE(capacity == 0 || (vals != nullptr && rows != nullptr && cols != nullptr));
    if (idx_bytes == 4) RB_REQUIRE(m <= 2147483647LL && n <= 2147483647LL);
    long long total = 0;
    RB_REQUIRE(!__builtin_mul_overflow((long long) m, (long long) n, &total));
    if (ambiguous) *ambiguous = 0;
    if (density == 0.0 || m == 0 || n == 0) {                // :302-304: nothing drawn
        *nnz = 0;
        store_ctr(load_ctr(ctr), next_ctr);
        return 0;
    }
    RB_REQUIRE(density < 1.0);   // density == 1 (:308-317, a dense matrix in COO form) is served by fill_dense, not here
    const Ctr128 c = load_ctr(ctr);
    const PhiloxKey k{key[0], key[1]};
    if (idx_bytes == 4) return random_coo_t<T, int32_t>(m, n, density, c, k, capacity, vals, (int32_t*) rows, (int32_t*) cols, nnz, next_ctr, ambiguous, st);
    return random_coo_t<T, int64_t>(m, n, density, c, k, capacity, vals, (int64_t*) rows, (int64_t*) cols, nnz, next_ctr, ambiguous, st);
}

template <typename T, typename IDX>
static int csr_colblock_t(int64_t n_rows, int64_t nnz, const T* vals, const IDX* rowptr, const IDX* colidxs, int64_t c0, int64_t c1,
                          int64_t capacity, T* out_vals, IDX* out_rowptr, IDX* out_cols, int64_t* nnz_out, cudaStream_t st) {
    unsigned long long* counts = (unsigned long long*) workspace(0, (size_t) (n_rows + 1) * 8, st);
    unsigned long long* optr = (unsigned long long*) workspace(1, (size_t) (n_rows + 1) * 8, st);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, optr, (int) (n_rows + 1), st);
    void* tmp = workspace(2, tmp_bytes, st);
    if (!counts || !optr || !tmp) return fail_cuda(cudaErrorMemoryAllocation, "column-block workspace");
    int64_t grid = (n_rows * 32 + 255) / 256;
    const int64_t cap = (int64_t) sm_count() * 16;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    colblock_count_kernel<IDX><<<(unsigned) grid, 256, 0, st>>>(n_rows, rowptr, colidxs, c0, c1, counts);
    RB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, optr, (int) (n_rows + 1), st));
    unsigned long long total = 0;
    RB_CUDA(cudaMemcpyAsync(&total, optr + n_rows, 8, cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    *nnz_out = (int64_t) total;
    const int write_entries = capacity >= (int64_t) total ? 1 : 0;
    colblock_write_kernel<T, IDX><<<(unsigned) grid, 256, 0, st>>>(n_rows, rowptr, colidxs, vals, c0, c1, optr, out_rowptr,
                                                                   out_cols, out_vals, write_entries);
    count_launch(3);
    RB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rb

using namespace rb;

extern "C" {

int rb_random_coo_f32(int64_t m, int64_t n, double density, const uint32_t ctr[4], const uint32_t key[2], int64_t capacity,
                      float* vals, void* rows, void* cols, int idx_bytes, int64_t* nnz, uint32_t next_ctr[4],
                      int64_t* ambiguous, void* stream) {
    return random_coo_impl<float>(m, n, density, ctr, key, capacity, vals, rows, cols, idx_bytes, nnz, next_ctr, ambiguous, stream);
}
int rb_random_coo_f64(int64_t m, int64_t n, double density, const uint32_t ctr[4], const uint32_t key[2], int64_t capacity,
                      double* vals, void* rows, void* cols, int idx_bytes, int64_t* nnz, uint32_t next_ctr[4],
                      int64_t* ambiguous, void* stream) {
    return random_coo_impl<double>(m, n, density, ctr, key, capacity, vals, rows, cols, idx_bytes, nnz, next_ctr, ambiguous, stream);
}

int rb_sorted_idxs_to_ptr(int64_t n_major, int64_t nnz, const void* idxs, int idx_bytes, void* ptr, void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(n_major >= 0 && nnz >= 0);
    RB_REQUIRE(idx_bytes == 4 || idx_bytes == 8);
    RB_REQUIRE(ptr != nullptr && (nnz == 0 || idxs != nullptr));
    int64_t grid = (n_major + 1 + 255) / 256;
    const int64_t cap = (int64_t) sm_count() * 16;
    if (grid > cap) grid = cap;
    if (idx_bytes == 4) sorted_to_ptr_kernel<int32_t><<<(unsigned) grid, 256, 0, st>>>(n_major, nnz, (const int32_t*) idxs, (int32_t*) ptr);
    else sorted_to_ptr_kernel<int64_t><<<(unsigned) grid, 256, 0, st>>>(n_major, nnz, (const int64_t*) idxs, (int64_t*) ptr);
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}

int rb_csr_column_block(int64_t n_rows, int64_t n_cols, int64_t nnz, const void* vals, int val_bytes, const void* rowptr,
                        const void* colidxs, int idx_bytes, int64_t c0, int64_t c1, int64_t capacity, void* out_vals,
                        void* out_rowptr, void* out_colidxs, int64_t* nnz_out, void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0 && capacity >= 0);
    RB_REQUIRE(0 <= c0 && c0 <= c1 && c1 <= n_cols);
    RB_REQUIRE(idx_bytes == 4 || idx_bytes == 8);
    RB_REQUIRE(val_bytes == 4 || val_bytes == 8);
    RB_REQUIRE(rowptr != nullptr && out_rowptr != nullptr && nnz_out != nullptr);
    RB_REQUIRE(nnz == 0 || (vals != nullptr && colidxs != nullptr));
    RB_REQUIRE(capacity == 0 || (out_vals != nullptr && out_colidxs != nullptr));
    RB_REQUIRE(n_rows + 1 <= 2147483647LL);
#define RB_CB(T, I) csr_colblock_t<T, I>(n_rows, nnz, (const T*) vals, (const I*) rowptr, (const I*) colidxs, c0, c1, capacity, \
                                         (T*) out_vals, (I*) out_rowptr, (I*) out_colidxs, nnz_out, st)
    if (val_bytes == 4) return idx_bytes == 4 ? RB_CB(float, int32_t) : RB_CB(float, int64_t);
    return idx_bytes == 4 ? RB_CB(double, int32_t) : RB_CB(double, int64_t);
#undef RB_CB
}

}  // extern "C"
