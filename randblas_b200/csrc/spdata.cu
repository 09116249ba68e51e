// K4: dense sketching operator applied to sparse data (sketch_sparse).
//
// Replaces sparse_data::lsksp3 / rsksp3 (RandBLAS/sparse_data/sksp.hh:132-182, 277-326) together with
// submatrix_as_blackbox (dense_skops.hh:677-688, which materialises the whole d x m operator on the host) and
// the right_spmm/left_spmm kernels (spmm_dispatch.hh:52-219; csc_spmm_impl.hh:99-209; csr_spmm_impl.hh:77-155).
//
// Canonical form: C(P x Q) = alpha * X(P x K) * Ysp(K x Q) + beta * C, X = op(S window), Ysp = op(A_sp window).
//
// Two kernels:
//   * spdata_kgroup_kernel (default): input-stationary. A warp owns four consecutive rows k of Ysp (a "k-group",
//     aligned so that with the operator's major axis along k the four rows share their Philox blocks) and 128 rows
//     i of C; each lane regenerates a 4(i) x 4(k) patch of X from (key, counter) ONCE -- 4 Philox blocks -- and then
//     walks the nonzeros (q, a) of those rows, adding a * X[i..i+3, k] into C[i..i+3, q] with one
//     red.global.add.v4.f32. The operator column is reused by every nonzero of the row instead of being
//     regenerated per nonzero, so the Philox/Box-Muller work is that of ONE pass over X (K*P samples) rather than
//     nnz*P blocks. The 128-row blocks of C are processed one after the other (outermost loop), so the slab of C
//     that receives the reductions (128 * Q * sizeof(T)) stays resident in the 126 MB L2 at the benchmark shape
//     and the reductions never go to HBM; the nonzeros are re-read once per slab (12 B each, negligible next to
//     the 512 B of reductions they cause). Row access to Ysp is free for CSR (not transposed) and CSC (transposed).
//     Floating-point reductions commute only up to rounding: results can differ in the last bits between runs
//     (well inside the rel-Frobenius contract); rb_set_option("spdata_path", 1) selects the kernel below instead.
//   * spdata_colowner_kernel: output-stationary, no atomics, bit-reproducible: one warp owns one column q of C,
//     keeps its P accumulators in registers and regenerates the operator column per nonzero (P Philox blocks per
//     nonzero when the operator's major axis runs along K). Column access to Ysp is free for CSC (not transposed)
//     and CSR (transposed).
// Format/transposition combinations that are not directly accessible the way the chosen kernel needs are first
// re-bucketed on the device (count, exclusive scan, scatter).
//
// Roofline: HBM on the bytes of A (12 B per nonzero with int64 indices) + C; DESIGN.md has the numbers.
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

// vectorised float reduction (sm_90+): one L2 transaction for four consecutive floats
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// the same with an L2 eviction-priority hint (evict_last: the slab of C that is being reduced into should stay in L2)
__device__ __forceinline__ unsigned long long l2_evict_last_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void red_add_v4_hint(float* addr, float a, float b, float c, float d, unsigned long long pol) {
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d),
                 "l"(pol)
                 : "memory");
}

// Input-stationary kernel: see the file header. Lane l of a warp owns rows i0..i0+3 of C, i0 = 128*ib + 4*l - ri.
// u_along_k = 1: X[i, k] = lane (u0+k)%4 of block (v0+i, (u0+k)/4); k-groups start at k = 4g - (u0 & 3).
// u_along_k = 0: X[i, k] = lane (u0+i)%4 of block (v0+k, (u0+i)/4); i-quads start at i = 4j - (u0 & 3) (ri).
template <typename T, typename IDX, bool GAUSS, bool VEC4>
__global__ void __launch_bounds__(256) spdata_kgroup_kernel(const SpDataProblem<T> p, const int64_t* __restrict__ ptr64,
                                                            const IDX* __restrict__ ptrN, const IDX* __restrict__ qidx,
                                                            const T* __restrict__ vals, int64_t seg_off, int64_t q_off,
                                                            int n_iblocks, int64_t n_groups, int u_along_k) {
    __shared__ __align__(16) double2 logtab[GAUSS ? LOGF_TABLE_ENTRIES : 1];
    if constexpr (GAUSS) { load_logf_table(logtab, p.gen.logtab); __syncthreads(); }
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    const int rk = u_along_k ? (int) (p.u0 & 3) : 0;
    const int ri = u_along_k ? 0 : (int) (p.u0 & 3);
    const unsigned long long pol = l2_evict_last_policy();
    for (int ib = 0; ib < n_iblocks; ++ib) {
        const int64_t i0 = (int64_t) ib * 128 + 4 * lane - ri;
        const bool i_live = (i0 + 3 >= 0) && (i0 < p.P);
        for (int64_t g = warp; g < n_groups; g += nwarps) {
            const int64_t k0 = 4 * g - rk;
            // segment boundaries of rows k0 .. k0+3 (clamped to the window [0, K))
            int64_t bnd = 0;
            if (lane < 5) {
                int64_t kk = k0 + lane;
                kk = kk < 0 ? 0 : (kk > p.K ? p.K : kk);
                bnd = ptr64 ? ptr64[seg_off + kk] : (int64_t) ptrN[seg_off + kk];
            }
            int64_t e[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) e[j] = __shfl_sync(0xffffffffu, bnd, j);
            if (e[4] == e[0]) continue;
            T s[4][4];   // s[ii][kk] = X[i0 + ii, k0 + kk]
            if (i_live) {
                if (u_along_k) {
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) {
                        const int64_t v = p.v0 + i0 + ii;
                        const uint4 w = philox4x32_10(ctr_add(p.gen.ctr, (uint64_t) (v * p.gen.R + ((p.u0 + k0) >> 2))), p.gen.key);
                        const float4 f = transform4<GAUSS>(w, logtab);
                        s[ii][0] = finish_sample<T, GAUSS>(f.x); s[ii][1] = finish_sample<T, GAUSS>(f.y);
                        s[ii][2] = finish_sample<T, GAUSS>(f.z); s[ii][3] = finish_sample<T, GAUSS>(f.w);
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        if (e[kk + 1] == e[kk]) continue;
                        const int64_t v = p.v0 + k0 + kk;
                        const uint4 w = philox4x32_10(ctr_add(p.gen.ctr, (uint64_t) (v * p.gen.R + ((p.u0 + i0) >> 2))), p.gen.key);
                        const float4 f = transform4<GAUSS>(w, logtab);
                        s[0][kk] = finish_sample<T, GAUSS>(f.x); s[1][kk] = finish_sample<T, GAUSS>(f.y);
                        s[2][kk] = finish_sample<T, GAUSS>(f.z); s[3][kk] = finish_sample<T, GAUSS>(f.w);
                    }
                }
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                for (int64_t eb = e[kk]; eb < e[kk + 1]; eb += 32) {
                    int my_q = -1;
                    T my_a = (T) 0;
                    if (eb + lane < e[kk + 1]) {
                        // streamed once per slab: evict-first, so that the slab of C stays resident in L2
                        const int64_t q = (int64_t) __ldcs(qidx + eb + lane) - q_off;
                        if (q >= 0 && q < p.Q) { my_q = (int) q; my_a = p.alpha * __ldcs(vals + eb + lane); }
                    }
                    const int cnt = (int) min((int64_t) 32, e[kk + 1] - eb);
                    for (int t = 0; t < cnt; ++t) {
                        const int q = __shfl_sync(0xffffffffu, my_q, t);
                        const T a = __shfl_sync(0xffffffffu, my_a, t);
                        if (q < 0 || !i_live) continue;      // outside the window of A_sp
                        T* c = p.C + (int64_t) q * p.ccs + i0 * p.crs;
                        if constexpr (VEC4) {
                            red_add_v4_hint((float*) c, a * s[0][kk], a * s[1][kk], a * s[2][kk], a * s[3][kk], pol);
                        } else {
#pragma unroll
                            for (int ii = 0; ii < 4; ++ii)
                                if (i0 + ii >= 0 && i0 + ii < p.P) atomicAdd(c + ii * p.crs, a * s[ii][kk]);
                        }
                    }
                }
            }
        }
    }
}

// ---- re-bucketing of op(A_sp window): by output column q (BYK = false) or by row k of Ysp (BYK = true) ----
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

template <typename T, typename IDX, bool BYK>
__global__ void __launch_bounds__(256) bucket_count_kernel(const SpDataProblem<T> p, unsigned long long* __restrict__ cnt) {
    NzIter<T, IDX>::run(p, [&](int64_t r, int64_t c, T) {
        const int64_t k = (p.y_is_transposed ? c - p.co_a : r - p.ro_a);
        const int64_t q = (p.y_is_transposed ? r - p.ro_a : c - p.co_a);
        if (k >= 0 && k < p.K && q >= 0 && q < p.Q) atomicAdd(cnt + (BYK ? k : q), 1ull);
    });
}

template <typename T, typename IDX, bool BYK>
__global__ void __launch_bounds__(256) bucket_scatter_kernel(const SpDataProblem<T> p, unsigned long long* __restrict__ cursor,
                                                             IDX* __restrict__ oidx, T* __restrict__ vals) {
    NzIter<T, IDX>::run(p, [&](int64_t r, int64_t c, T a) {
        const int64_t k = (p.y_is_transposed ? c - p.co_a : r - p.ro_a);
        const int64_t q = (p.y_is_transposed ? r - p.ro_a : c - p.co_a);
        if (k >= 0 && k < p.K && q >= 0 && q < p.Q) {
            const unsigned long long pos = atomicAdd(cursor + (BYK ? k : q), 1ull);
            oidx[pos] = (IDX) (BYK ? q : k);
            vals[pos] = a;
        }
    });
}

// Buckets the window of op(A_sp) by q (BYK = false) or k (BYK = true) into workspace slots 1 (segment pointers,
// int64), 3 (the other index, window-relative) and 4 (values).
template <typename T, typename IDX, bool BYK>
int rebucket(const SpDataProblem<T>& p, cudaStream_t st, const int64_t** ptr_out, const IDX** oidx_out, const T** vals_out) {
    const int64_t nseg = (BYK ? p.K : p.Q) + 1;
    if (nseg > 2147483647LL) return fail("sketch_sparse: re-bucketing more than 2^31-2 segments is not supported");
    unsigned long long* cnt = (unsigned long long*) workspace(0, (size_t) nseg * 8, st);
    unsigned long long* ptr = (unsigned long long*) workspace(1, (size_t) nseg * 8, st);
    unsigned long long* cur = (unsigned long long*) workspace(2, (size_t) nseg * 8, st);
    IDX* bk = (IDX*) workspace(3, (size_t) (p.nnz > 0 ? p.nnz : 1) * sizeof(IDX), st);
    T* bv = (T*) workspace(4, (size_t) (p.nnz > 0 ? p.nnz : 1) * sizeof(T), st);
    if (!cnt || !ptr || !cur || !bk || !bv) return fail_cuda(cudaErrorMemoryAllocation, "sketch_sparse bucket workspace");
    RB_CUDA(cudaMemsetAsync(cnt, 0, (size_t) nseg * 8, st));
    int64_t units = (p.fmt == 0) ? p.A_rows : (p.fmt == 1 ? p.A_cols : (p.nnz + 31) / 32);
    int64_t grid = (units + 7) / 8;
    int64_t cap = (int64_t) sm_count() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    bucket_count_kernel<T, IDX, BYK><<<(unsigned) grid, 256, 0, st>>>(p, cnt);
    count_launch();
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, ptr, (int) nseg, st);
    void* tmp = workspace(5, tmp_bytes, st);
    if (!tmp) return fail_cuda(cudaErrorMemoryAllocation, "scan workspace");
    RB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, ptr, (int) nseg, st));
    count_launch();
    RB_CUDA(cudaMemcpyAsync(cur, ptr, (size_t) nseg * 8, cudaMemcpyDeviceToDevice, st));
    bucket_scatter_kernel<T, IDX, BYK><<<(unsigned) grid, 256, 0, st>>>(p, cur, bk, bv);
    count_launch();
    RB_CUDA(cudaGetLastError());
    *ptr_out = (const int64_t*) ptr;
    *oidx_out = bk;
    *vals_out = bv;
    return 0;
}

// input-stationary path
template <typename T, typename IDX>
int launch_spdata_kgroup(const SpDataProblem<T>& p, cudaStream_t st) {
    const bool direct = (p.fmt == 0 && !p.y_is_transposed) || (p.fmt == 1 && p.y_is_transposed);
    const int64_t* ptr64 = nullptr;
    const IDX* ptrN = nullptr;
    const IDX* qidx = nullptr;
    const T* vals = nullptr;
    int64_t seg_off = 0, q_off = 0;
    if (direct) {
        // CSR: segments are rows of A (k = row - ro_a), entries are column indices (q = col - co_a)
        // CSC^T: segments are columns of A (k = col - co_a), entries are row indices (q = row - ro_a)
        ptrN = (const IDX*) (p.fmt == 0 ? p.idx0 : p.idx1);
        qidx = (const IDX*) (p.fmt == 0 ? p.idx1 : p.idx0);
        vals = p.vals;
        seg_off = (p.fmt == 0) ? p.ro_a : p.co_a;
        q_off = (p.fmt == 0) ? p.co_a : p.ro_a;
    } else {
        int rc = rebucket<T, IDX, true>(p, st, &ptr64, &qidx, &vals);
        if (rc) return rc;
    }
    int rc = launch_scale<T>(p.P, p.Q, p.beta, p.C, p.crs, p.ccs, st);
    if (rc) return rc;
    const int u_along_k = p.uk == 1 ? 1 : 0;
    const int rk = u_along_k ? (int) (p.u0 & 3) : 0, ri = u_along_k ? 0 : (int) (p.u0 & 3);
    const int64_t n_groups = (p.K + rk + 3) / 4;
    const int n_iblocks = (int) ((p.P + ri + 127) / 128);
    const bool vec4 = sizeof(T) == 4 && p.crs == 1 && (p.ccs & 3) == 0 && (p.P & 3) == 0 && ri == 0 &&
                      (reinterpret_cast<uintptr_t>(p.C) & 15) == 0;
    int64_t grid = (n_groups + 7) / 8;
    int64_t cap = (int64_t) sm_count() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    const bool gauss = p.family == 'G';
#define RB_KG(G, V) spdata_kgroup_kernel<T, IDX, G, V><<<(unsigned) grid, 256, 0, st>>>(p, ptr64, ptrN, qidx, vals, seg_off, q_off, n_iblocks, n_groups, u_along_k)
    if constexpr (sizeof(T) == 4) {
        if (vec4) { if (gauss) RB_KG(true, true); else RB_KG(false, true); }
        else { if (gauss) RB_KG(true, false); else RB_KG(false, false); }
    } else {
        if (gauss) RB_KG(true, false); else RB_KG(false, false);
    }
#undef RB_KG
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
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
        int rc = rebucket<T, IDX, false>(p, st, &ptr64, &kidx, &vals);
        if (rc) return rc;
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
    if (p.Q > 2147483647LL) return fail("sketch_sparse: more than 2^31-1 output columns is not supported");
    if (p.family == 'G' && !p.gen.logtab) return fail_cuda(cudaErrorMemoryAllocation, "logf table of the Gaussian generator");
    if (get_option("spdata_path") == 1) {
        if (p.idx_bytes == 4) return launch_spdata_t<T, int32_t>(p, st);
        return launch_spdata_t<T, int64_t>(p, st);
    }
    if (p.idx_bytes == 4) return launch_spdata_kgroup<T, int32_t>(p, st);
    return launch_spdata_kgroup<T, int64_t>(p, st);
}
template int launch_spdata<float>(const SpDataProblem<float>&, cudaStream_t);
template int launch_spdata<double>(const SpDataProblem<double>&, cudaStream_t);

}  // namespace rb
