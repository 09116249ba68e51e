// K4: dense sketching operator applied to sparse data (sketch_sparse).
//
// Replaces sparse_data::lsksp3 / rsksp3 (RandBLAS/sparse_data/sksp.hh:132-182, 277-326) together with
// submatrix_as_blackbox (dense_skops.hh:677-688, which materialises the whole d x m operator on the host) and
// the right_spmm/left_spmm kernels (spmm_dispatch.hh:52-219; csc_spmm_impl.hh:99-209; csr_spmm_impl.hh:77-155).
//
// Canonical form: C(P x Q) = alpha * X(P x K) * Ysp(K x Q) + beta * C, X = op(S window), Ysp = op(A_sp window).
// The kernel is output-stationary: one warp owns one column q of C, keeps its P accumulators in registers,
// walks the nonzeros (k, a) of column q of Ysp and regenerates the operator column X[:, k] on the fly from
// (key, counter) -- S never exists in memory and C is written exactly once (no atomics, no read-modify-write
// of the d-long output columns per nonzero, which is what makes the CPU kernel's axpy formulation
// memory-bound). Column access to Ysp is free for CSC (not transposed) and CSR (transposed); the other
// format/transposition combinations are first re-bucketed by output column on the device (count, exclusive
// scan, scatter).
//
// Roofline: at the benchmark shape the kernel is bound by integer/FP issue for regenerating X (P Philox
// blocks per nonzero when the operator's major axis runs along K), not by HBM; DESIGN.md has the numbers.
#include <cub/device/device_scan.cuh>
#include "common.cuh"
#include "kernels.h"

namespace rb {

namespace {

// one operator entry from natural coordinates (v, u)
template <typename T, bool GAUSS>
__device__ __forceinline__ T gen_entry(const DenseGen& g, int64_t v, int64_t u, const double2* logtab) {
    const uint4 w = philox4x32_10(ctr_add(g.ctr, (uint64_t) (v * g.R + (u >> 2))), g.key);
    const int lane = (int) (u & 3);
    float f;
    if constexpr (GAUSS) {
        float g0, g1;
        if (lane < 2) boxmuller(w.x, w.y, logtab, g0, g1); else boxmuller(w.z, w.w, logtab, g0, g1);
        f = (lane & 1) ? g1 : g0;
    } else {
        const uint32_t ww = lane == 0 ? w.x : (lane == 1 ? w.y : (lane == 2 ? w.z : w.w));
        f = uneg11f(ww);
    }
    return finish_sample<T, GAUSS>(f);
}

// ACC accumulators per lane: rows i = pbase + lane + 32*t  (t < ACC) when the operator's u axis runs along k,
// or rows i = pbase + 4*(lane + 32*t) + {0..3} (ACC multiple of 4) when u runs along i and is 4-aligned.
template <typename T, typename IDX, bool GAUSS, int ACC>
__global__ void __launch_bounds__(256) spdata_colowner_kernel(const SpDataProblem<T> p, const int64_t* __restrict__ ptr64,
                                                              const IDX* __restrict__ ptrN, const IDX* __restrict__ kidx,
                                                              const T* __restrict__ vals, int64_t seg_off,
                                                              int64_t k_off, int u_blocked) {
    __shared__ __align__(16) double2 logtab[GAUSS ? LOGF_TABLE_ENTRIES : 1];
    if constexpr (GAUSS) { load_logf_table(logtab, p.gen.logtab); __syncthreads(); }
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    for (int64_t q = warp; q < p.Q; q += nwarps) {
        const int64_t e_lo = ptr64 ? ptr64[seg_off + q] : (int64_t) ptrN[seg_off + q];
        const int64_t e_hi = ptr64 ? ptr64[seg_off + q + 1] : (int64_t) ptrN[seg_off + q + 1];
        for (int64_t pbase = 0; pbase < p.P; pbase += 32 * ACC) {
            T acc[ACC];
#pragma unroll
            for (int t = 0; t < ACC; ++t) acc[t] = (T) 0;
            for (int64_t eb = e_lo; eb < e_hi; eb += 32) {
                int64_t my_k = -1;
                T my_a = (T) 0;
                if (eb + lane < e_hi) { my_k = (int64_t) kidx[eb + lane] - k_off; my_a = vals[eb + lane]; }
                const int cnt = (int) min((int64_t) 32, e_hi - eb);
                for (int s = 0; s < cnt; ++s) {
                    const int64_t k = __shfl_sync(0xffffffffu, my_k, s);
                    const T a = __shfl_sync(0xffffffffu, my_a, s);
                    if (k < 0 || k >= p.K) continue;          // outside the window of A_sp
                    if (u_blocked) {
#pragma unroll
                        for (int t = 0; t < ACC; t += 4) {
                            const int64_t i = pbase + 4 * (lane + 32 * (t / 4));
                            if (i < p.P) {
                                const int64_t v = p.v0 + k;       // vk == 1, ui == 1
                                const int64_t u = p.u0 + i;       // multiple of 4 by construction
                                const uint4 w = philox4x32_10(ctr_add(p.gen.ctr, (uint64_t) (v * p.gen.R + (u >> 2))), p.gen.key);
                                const float4 f = transform4<GAUSS>(w, logtab);
                                acc[t + 0] += a * finish_sample<T, GAUSS>(f.x);
                                if (i + 1 < p.P) acc[t + 1] += a * finish_sample<T, GAUSS>(f.y);
                                if (i + 2 < p.P) acc[t + 2] += a * finish_sample<T, GAUSS>(f.z);
                                if (i + 3 < p.P) acc[t + 3] += a * finish_sample<T, GAUSS>(f.w);
                            }
                        }
                    } else {
#pragma unroll
                        for (int t = 0; t < ACC; ++t) {
                            const int64_t i = pbase + lane + 32 * t;
                            if (i < p.P) {
                                const int64_t v = p.v0 + i * p.vi + k * p.vk;
                                const int64_t u = p.u0 + i * p.ui + k * p.uk;
                                acc[t] += a * gen_entry<T, GAUSS>(p.gen, v, u, logtab);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < ACC; ++t) {
                const int64_t i = u_blocked ? (pbase + 4 * (lane + 32 * (t / 4)) + (t & 3)) : (pbase + lane + 32 * t);
                if (i < p.P) {
                    T* c = p.C + i * p.crs + q * p.ccs;
                    T r = p.alpha * acc[t];
                    if (p.beta != (T) 0) r += p.beta * (*c);
                    *c = r;
                }
            }
        }
    }
}

// ---- re-bucketing of op(A_sp window) by output column q ----
template <typename T, typename IDX>
struct NzIter {
    // calls f(r, c, val) for every stored nonzero, warp-cooperatively; (r, c) are coordinates in A_sp
    template <typename F>
    static __device__ __forceinline__ void run(const SpDataProblem<T>& p, F f) {
        const IDX* i0 = (const IDX*) p.idx0;
        const IDX* i1 = (const IDX*) p.idx1;
        const int lane = threadIdx.x & 31;
        const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const int64_t nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
        if (p.fmt == 0) {          // CSR: idx0 = rowptr, idx1 = colidxs
            for (int64_t r = warp; r < p.A_rows; r += nwarps)
                for (int64_t e = (int64_t) i0[r] + lane; e < (int64_t) i0[r + 1]; e += 32) f(r, (int64_t) i1[e], p.vals[e]);
        } else if (p.fmt == 1) {   // CSC: idx0 = rowidxs, idx1 = colptr
            for (int64_t c = warp; c < p.A_cols; c += nwarps)
                for (int64_t e = (int64_t) i1[c] + lane; e < (int64_t) i1[c + 1]; e += 32) f((int64_t) i0[e], c, p.vals[e]);
        } else {                   // COO
            for (int64_t e = warp * 32 + lane; e < p.nnz; e += nwarps * 32) f((int64_t) i0[e], (int64_t) i1[e], p.vals[e]);
        }
    }
};

template <typename T, typename IDX>
__global__ void __launch_bounds__(256) bucket_count_kernel(const SpDataProblem<T> p, unsigned long long* __restrict__ cnt) {
    NzIter<T, IDX>::run(p, [&](int64_t r, int64_t c, T) {
        const int64_t k = (p.y_is_transposed ? c - p.co_a : r - p.ro_a);
        const int64_t q = (p.y_is_transposed ? r - p.ro_a : c - p.co_a);
        if (k >= 0 && k < p.K && q >= 0 && q < p.Q) atomicAdd(cnt + q, 1ull);
    });
}

template <typename T, typename IDX>
__global__ void __launch_bounds__(256) bucket_scatter_kernel(const SpDataProblem<T> p, unsigned long long* __restrict__ cursor,
                                                             IDX* __restrict__ kidx, T* __restrict__ vals) {
    NzIter<T, IDX>::run(p, [&](int64_t r, int64_t c, T a) {
        const int64_t k = (p.y_is_transposed ? c - p.co_a : r - p.ro_a);
        const int64_t q = (p.y_is_transposed ? r - p.ro_a : c - p.co_a);
        if (k >= 0 && k < p.K && q >= 0 && q < p.Q) {
            const unsigned long long pos = atomicAdd(cursor + q, 1ull);
            kidx[pos] = (IDX) k;
            vals[pos] = a;
        }
    });
}

template <typename T, typename IDX>
int launch_spdata_t(const SpDataProblem<T>& p, cudaStream_t st) {
    const bool direct = (p.fmt == 1 && !p.y_is_transposed) || (p.fmt == 0 && p.y_is_transposed);
    const int64_t* ptr64 = nullptr;
    const IDX* ptrN = nullptr;
    const IDX* kidx = nullptr;
    const T* vals = nullptr;
    int64_t seg_off = 0, k_off = 0;
    if (direct) {
        // CSC: segments are columns of A (q = col - co_a), entries are row indices (k = row - ro_a)
        // CSR^T: segments are rows of A (q = row - ro_a), entries are column indices (k = col - co_a)
        ptrN = (const IDX*) (p.fmt == 1 ? p.idx1 : p.idx0);
        kidx = (const IDX*) (p.fmt == 1 ? p.idx0 : p.idx1);
        vals = p.vals;
        seg_off = (p.fmt == 1) ? p.co_a : p.ro_a;
        k_off = (p.fmt == 1) ? p.ro_a : p.co_a;
    } else {
        const int64_t nseg = p.Q + 1;
        unsigned long long* cnt = (unsigned long long*) workspace(0, (size_t) nseg * 8);
        unsigned long long* ptr = (unsigned long long*) workspace(1, (size_t) nseg * 8);
        unsigned long long* cur = (unsigned long long*) workspace(2, (size_t) nseg * 8);
        IDX* bk = (IDX*) workspace(3, (size_t) (p.nnz > 0 ? p.nnz : 1) * sizeof(IDX));
        T* bv = (T*) workspace(4, (size_t) (p.nnz > 0 ? p.nnz : 1) * sizeof(T));
        if (!cnt || !ptr || !cur || !bk || !bv) return fail_cuda(cudaErrorMemoryAllocation, "sketch_sparse bucket workspace");
        RB_CUDA(cudaMemsetAsync(cnt, 0, (size_t) nseg * 8, st));
        int64_t units = (p.fmt == 0) ? p.A_rows : (p.fmt == 1 ? p.A_cols : (p.nnz + 31) / 32);
        int64_t grid = (units + 7) / 8;
        int64_t cap = (int64_t) sm_count() * 8;
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        bucket_count_kernel<T, IDX><<<(unsigned) grid, 256, 0, st>>>(p, cnt);
        count_launch();
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, ptr, (int) nseg, st);
        void* tmp = workspace(5, tmp_bytes);
        if (!tmp) return fail_cuda(cudaErrorMemoryAllocation, "scan workspace");
        RB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, ptr, (int) nseg, st));
        count_launch();
        RB_CUDA(cudaMemcpyAsync(cur, ptr, (size_t) nseg * 8, cudaMemcpyDeviceToDevice, st));
        bucket_scatter_kernel<T, IDX><<<(unsigned) grid, 256, 0, st>>>(p, cur, bk, bv);
        count_launch();
        RB_CUDA(cudaGetLastError());
        ptr64 = (const int64_t*) ptr;
        kidx = bk;
        vals = bv;
    }
    // u runs along i (blocks shared by 4 consecutive output rows) and is 4-aligned?
    const int u_blocked = (p.ui == 1 && p.vk == 1 && (p.u0 & 3) == 0) ? 1 : 0;
    int64_t grid = (p.Q + 7) / 8;
    int64_t cap = (int64_t) sm_count() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    const bool gauss = p.family == 'G';
#define RB_SPD(G, ACC) spdata_colowner_kernel<T, IDX, G, ACC><<<(unsigned) grid, 256, 0, st>>>(p, ptr64, ptrN, kidx, vals, seg_off, k_off, u_blocked)
    if (p.P <= 128) { if (gauss) RB_SPD(true, 4); else RB_SPD(false, 4); }
    else if (p.P <= 256) { if (gauss) RB_SPD(true, 8); else RB_SPD(false, 8); }
    else { if (gauss) RB_SPD(true, 16); else RB_SPD(false, 16); }
#undef RB_SPD
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

template <typename T>
int launch_spdata(const SpDataProblem<T>& p, cudaStream_t st) {
    if (p.P <= 0 || p.Q <= 0) return 0;
    if (p.K <= 0 || p.alpha == (T) 0 || p.nnz <= 0) return launch_scale<T>(p.P, p.Q, p.beta, p.C, p.crs, p.ccs, st);
    if (p.Q + 1 > 2147483647LL) return fail("sketch_sparse: more than 2^31-2 output columns is not supported");
    if (p.idx_bytes == 4) return launch_spdata_t<T, int32_t>(p, st);
    return launch_spdata_t<T, int64_t>(p, st);
}
template int launch_spdata<float>(const SpDataProblem<float>&, cudaStream_t);
template int launch_spdata<double>(const SpDataProblem<double>&, cudaStream_t);

}  // namespace rb
