// K3: short-axis-sparse operators (SASO).
//
//  * saso_fill_kernel   -- replaces sparse::repeated_fisher_yates (RandBLAS/sparse_skops.hh:51-106) and the SASO
//                          branch of fill_sparse_unpacked_nosub (:515-533). The reference is a serial loop that
//                          restores its work permutation after every vector (:92-102), so every minor-axis
//                          vector is an independent Fisher-Yates run over a (virtual) identity permutation at
//                          counters seed + i*vec_nnz + j. One lane per step computes the pivot; the value
//                          sitting at the pivot position is recovered by tracing the earlier swaps backwards
//                          (O(k^2) register work, no work array). Output goes through shared memory so that
//                          the three COO arrays are written with full-width coalesced stores.
//  * saso_apply_kernel  -- replaces sparse::lskges / rskges (RandBLAS/skge.hh:465-492, 598-626) and the COO->CSC
//                          sort + CSC kernels under left_spmm (sparse_data/coo_spmm_impl.hh:53-105,
//                          csc_spmm_impl.hh:99-209): the (index, sign) list of each vector is regenerated in
//                          registers, the operator's COO arrays are never materialised, sorted or re-read.
//  * coo_apply_kernel   -- the same contraction for an already-sampled operator handed over as COO arrays.
//
// Roofline: HBM. fill: bytes written = full_nnz * (2*idx_bytes + val_bytes). apply: bytes of A read once.
#include "common.cuh"
#include "kernels.h"

namespace rb {

namespace {

// pivot of Fisher-Yates step j: j + w0 % (dim_major - j)   (sparse_skops.hh:78; unsigned 32 -> 64 bit modulo)
__device__ __forceinline__ int64_t fy_pivot(uint32_t w0, int64_t j, int64_t dim_major) {
    const uint64_t rem = (uint64_t) (dim_major - j);
    const uint64_t r = (rem > 0xffffffffull) ? (uint64_t) w0 : (uint64_t) (w0 % (uint32_t) rem);
    return j + (int64_t) r;
}

// One warp computes one vector of k <= 32 entries: lane j holds step j's (major index, sign).
// The value at position `ell` after swaps 0..j-1 of an identity permutation: walk the swaps backwards. Swap t
// exchanges positions t and piv_t >= t. The traced position starts at piv_j >= j > t and, once it has been moved, sits
// at an earlier step's index t' > t, so it can never equal t itself: the only case left per step is
// "pos == piv_t -> pos = t" (one compare and one select instead of three and two).
__device__ __forceinline__ void saso_vector_warp(const Ctr128& base, const PhiloxKey& key, int64_t vec, int k,
                                                 int64_t dim_major, int lane, int64_t& major, int& negative) {
    int64_t piv = 0;
    uint32_t w1 = 0;
    if (lane < k) {
        const uint4 w = philox4x32_10(ctr_add(base, (uint64_t) (vec * k + lane)), key);
        piv = fy_pivot(w.x, lane, dim_major);
        w1 = w.y;
    }
    int64_t pos = piv;
    for (int t = k - 2; t >= 0; --t) {
        const int64_t pt = __shfl_sync(0xffffffffu, piv, t);
        if (t < lane && pos == pt) pos = t;
    }
    major = pos;
    negative = (int) (w1 & 1u);
}

// General k (any size): one thread per vector, pivots kept in a global scratch row of length k.
__device__ __forceinline__ int64_t fy_trace(const int64_t* piv, int j, int64_t pos) {
    for (int t = j - 1; t >= 0; --t) {
        if (pos == piv[t]) pos = t;
    }
    return pos;
}

template <typename IDX, typename VAL>
__device__ __forceinline__ void put_entry(IDX* maj, IDX* mnr, VAL* vals, int64_t e, int64_t major, int64_t minor,
                                          int negative) {
    maj[e] = (IDX) major;
    if (mnr) mnr[e] = (IDX) minor;
    if (vals) vals[e] = negative ? (VAL) -1 : (VAL) 1;
}

// k <= 32: warp per vector, 8 warps per CTA, CTA-level staging so global stores are contiguous per CTA.
template <typename IDX, typename VAL>
__global__ void __launch_bounds__(256) saso_fill_warp_kernel(Ctr128 ctr, PhiloxKey key, int k, int64_t dim_major,
                                                             int64_t dim_minor, IDX* __restrict__ maj,
                                                             IDX* __restrict__ mnr, VAL* __restrict__ vals) {
    constexpr int VPC = 64;                 // vectors per CTA iteration (8 warps x 8 vectors)
    __shared__ int64_t s_major[VPC * 32];
    __shared__ unsigned char s_neg[VPC * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base_vec = (int64_t) blockIdx.x * VPC; base_vec < dim_minor; base_vec += (int64_t) gridDim.x * VPC) {
        const int nvec = (int) min((int64_t) VPC, dim_minor - base_vec);
        for (int vl = warp; vl < nvec; vl += 8) {
            int64_t major; int neg;
            saso_vector_warp(ctr, key, base_vec + vl, k, dim_major, lane, major, neg);
            if (lane < k) { s_major[vl * k + lane] = major; s_neg[vl * k + lane] = (unsigned char) neg; }
        }
        __syncthreads();
        const int64_t e0 = base_vec * k;
        for (int e = threadIdx.x; e < nvec * k; e += 256)
            put_entry<IDX, VAL>(maj, mnr, vals, e0 + e, s_major[e], base_vec + e / k, s_neg[e]);
        __syncthreads();
    }
}

// k <= 32 and dim_major < 2^31: G = 2^ceil(log2 k) lanes per vector, 32 / G vectors per warp (the warp-per-vector
// kernel above keeps 32 - k lanes idle: 75% of them at vec_nnz = 8). Lane (vector, j) writes entry vector * k + j
// directly: consecutive lanes hit consecutive addresses, so the stores are coalesced without staging.
template <typename IDX, typename VAL, int G>
__global__ void __launch_bounds__(256) saso_fill_group_kernel(Ctr128 ctr, PhiloxKey key, int k, uint32_t dim_major,
                                                              int64_t dim_minor, IDX* __restrict__ maj,
                                                              IDX* __restrict__ mnr, VAL* __restrict__ vals) {
    constexpr int VPW = 32 / G;
    const int lane = threadIdx.x & 31, sub = lane & (G - 1);
    const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    // the divisor of step `sub` is the same for every vector this lane works on
    const uint32_t dv = sub < k ? dim_major - (uint32_t) sub : 1u, dm = fastmod_magic(dv);
    for (int64_t wv = warp * VPW; wv < dim_minor; wv += nwarps * VPW) {
        const int64_t v = wv + lane / G;
        const bool live = v < dim_minor && sub < k;
        uint32_t piv = 0, w1 = 0;
        if (live) {
            const uint4 w = philox4x32_10(ctr_add(ctr, (uint64_t) (v * k + sub)), key);
            piv = (uint32_t) sub + fastmod(w.x, dv, dm);                    // sparse_skops.hh:78
            w1 = w.y;
        }
        // value at position piv after swaps 0..sub-1 of an identity permutation: walk the swaps backwards
        // (steps t >= sub do not apply to this lane, so the walk may start at G - 2 >= k - 2 and unroll)
        uint32_t pos = piv;
#pragma unroll
        for (int t = G - 2; t >= 0; --t) {
            const uint32_t pt = __shfl_sync(0xffffffffu, piv, t, G);
            if (t < sub && pos == pt) pos = (uint32_t) t;
        }
        if (live) put_entry<IDX, VAL>(maj, mnr, vals, v * k + sub, (int64_t) pos, v, (int) (w1 & 1u));
    }
}

// any k: thread per vector, pivots in global scratch (k entries per thread of the grid)
template <typename IDX, typename VAL>
__global__ void saso_fill_thread_kernel(Ctr128 ctr, PhiloxKey key, int64_t k, int64_t dim_major, int64_t dim_minor,
                                        IDX* __restrict__ maj, IDX* __restrict__ mnr, VAL* __restrict__ vals,
                                        int64_t* __restrict__ scratch) {
    const int64_t gtid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    int64_t* piv = scratch + gtid * k;
    for (int64_t vec = gtid; vec < dim_minor; vec += (int64_t) gridDim.x * blockDim.x) {
        for (int64_t j = 0; j < k; ++j) {
            const uint4 w = philox4x32_10(ctr_add(ctr, (uint64_t) (vec * k + j)), key);
            const int64_t p = fy_pivot(w.x, j, dim_major);
            piv[j] = p;
            put_entry<IDX, VAL>(maj, mnr, vals, vec * k + j, fy_trace(piv, (int) j, p), vec, (int) (w.y & 1u));
        }
    }
}

template <typename IDX, typename VAL>
int launch_saso_t(Ctr128 ctr, PhiloxKey key, int64_t k, int64_t dim_major, int64_t dim_minor, void* maj, void* mnr,
                  void* vals, cudaStream_t st) {
    if (k <= 32 && dim_major < 0x7fffffffLL && get_option("saso_fill_path") == 0) {
        const int G = k <= 1 ? 1 : k <= 2 ? 2 : k <= 4 ? 4 : k <= 8 ? 8 : k <= 16 ? 16 : 32;
        const int64_t vpb = 8 * (32 / G);                      // vectors per 256-thread CTA and iteration
        int64_t grid = (dim_minor + vpb - 1) / vpb;
        const int64_t cap = (int64_t) sm_count() * 16;
        if (grid > cap) grid = cap;
#define RB_SASO_GROUP(GG) saso_fill_group_kernel<IDX, VAL, GG><<<(unsigned) grid, 256, 0, st>>>(                       \
            ctr, key, (int) k, (uint32_t) dim_major, dim_minor, (IDX*) maj, (IDX*) mnr, (VAL*) vals)
        switch (G) {
            case 1: RB_SASO_GROUP(1); break;
            case 2: RB_SASO_GROUP(2); break;
            case 4: RB_SASO_GROUP(4); break;
            case 8: RB_SASO_GROUP(8); break;
            case 16: RB_SASO_GROUP(16); break;
            default: RB_SASO_GROUP(32); break;
        }
#undef RB_SASO_GROUP
    } else if (k <= 32) {
        int64_t grid = (dim_minor + 63) / 64;
        int64_t cap = (int64_t) sm_count() * 8;
        if (grid > cap) grid = cap;
        saso_fill_warp_kernel<IDX, VAL><<<(unsigned) grid, 256, 0, st>>>(ctr, key, (int) k, dim_major, dim_minor,
                                                                         (IDX*) maj, (IDX*) mnr, (VAL*) vals);
    } else {
        int64_t threads = 128;
        int64_t grid = (dim_minor + threads - 1) / threads;
        int64_t cap = (int64_t) sm_count() * 4;
        if (grid > cap) grid = cap;
        int64_t* scratch = (int64_t*) workspace(7, (size_t) (grid * threads * k) * sizeof(int64_t), st);
        if (!scratch) return fail_cuda(cudaErrorMemoryAllocation, "workspace for Fisher-Yates pivots");
        saso_fill_thread_kernel<IDX, VAL><<<(unsigned) grid, (unsigned) threads, 0, st>>>(
            ctr, key, k, dim_major, dim_minor, (IDX*) maj, (IDX*) mnr, (VAL*) vals, scratch);
    }
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// apply: every warp takes one minor-axis vector, regenerates its (major index, sign) list and
// adds alpha * sign * Y[k, :] into C[i, :]. C must have been beta-scaled already.
template <typename T>
__device__ __forceinline__ void row_axpy_atomic(T* __restrict__ c, int64_t ccs, const T* __restrict__ y, int64_t ycs,
                                                int64_t Q, T a, int lane) {
    if constexpr (sizeof(T) == 4) {
        if (ccs == 1 && ycs == 1 && ((reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
            const int64_t q4 = Q >> 2;
            for (int64_t q = lane; q < q4; q += 32) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(y) + q);
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(c + 4 * q), "f"(a * v.x),
                             "f"(a * v.y), "f"(a * v.z), "f"(a * v.w)
                             : "memory");
            }
            for (int64_t q = (q4 << 2) + lane; q < Q; q += 32) atomicAdd(c + q, a * y[q]);
            return;
        }
    }
    for (int64_t q = lane; q < Q; q += 32) atomicAdd(c + q * ccs, a * y[q * ycs]);
}

template <typename T>
__global__ void __launch_bounds__(256) saso_apply_kernel(const SasoProblem<T> p, int64_t vec_lo, int64_t vec_hi) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    const int k = (int) p.vec_nnz;
    for (int64_t vec = vec_lo + warp; vec < vec_hi; vec += nwarps) {
        int64_t major; int neg;
        saso_vector_warp(p.ctr, p.key, vec, k, p.dim_major, lane, major, neg);
        for (int j = 0; j < k; ++j) {
            const int64_t mj = __shfl_sync(0xffffffffu, major, j);
            const int ng = __shfl_sync(0xffffffffu, neg, j);
            // entry position in S coordinates
            const int64_t row = (p.major_is_rows ? mj : vec) - p.ro_s;
            const int64_t col = (p.major_is_rows ? vec : mj) - p.co_s;
            if (row < 0 || row >= p.rs || col < 0 || col >= p.cs) continue;
            const int64_t i = p.x_is_transposed ? col : row;
            const int64_t kk = p.x_is_transposed ? row : col;
            const T a = ng ? -p.alpha : p.alpha;
            row_axpy_atomic<T>(p.C + i * p.crs, p.ccs, p.Y + kk * p.yrs, p.ycs, p.Q, a, lane);
        }
    }
}

template <typename T, typename IDX>
__global__ void __launch_bounds__(256) coo_apply_kernel(const CooProblem<T> p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    const IDX* rows = (const IDX*) p.rows;
    const IDX* cols = (const IDX*) p.cols;
    for (int64_t e = warp; e < p.nnz; e += nwarps) {
        const int64_t row = (int64_t) rows[e] - p.ro_s, col = (int64_t) cols[e] - p.co_s;
        if (row < 0 || row >= p.rs || col < 0 || col >= p.cs) continue;
        const int64_t i = p.x_is_transposed ? col : row;
        const int64_t kk = p.x_is_transposed ? row : col;
        row_axpy_atomic<T>(p.C + i * p.crs, p.ccs, p.Y + kk * p.yrs, p.ycs, p.Q, p.alpha * p.vals[e], lane);
    }
}

}  // namespace

int launch_saso(Ctr128 ctr, PhiloxKey key, int64_t vec_nnz, int64_t dim_major, int64_t dim_minor, void* idxs_major,
                void* idxs_minor, int idx_bytes, void* vals, int val_bytes, cudaStream_t st) {
    if (dim_minor <= 0 || vec_nnz <= 0) return 0;
    if (idx_bytes == 4) {
        if (val_bytes == 4) return launch_saso_t<int32_t, float>(ctr, key, vec_nnz, dim_major, dim_minor, idxs_major, idxs_minor, vals, st);
        return launch_saso_t<int32_t, double>(ctr, key, vec_nnz, dim_major, dim_minor, idxs_major, idxs_minor, vals, st);
    }
    if (val_bytes == 4) return launch_saso_t<int64_t, float>(ctr, key, vec_nnz, dim_major, dim_minor, idxs_major, idxs_minor, vals, st);
    return launch_saso_t<int64_t, double>(ctr, key, vec_nnz, dim_major, dim_minor, idxs_major, idxs_minor, vals, st);
}

template <typename T>
int launch_saso_apply(const SasoProblem<T>& p, cudaStream_t st) {
    if (p.P <= 0 || p.Q <= 0) return 0;
    if (p.vec_nnz > 32 && p.K > 0 && p.alpha != (T) 0)
        return -1;   // caller materialises the COO arrays and uses launch_coo_apply
    int rc = launch_scale<T>(p.P, p.Q, p.beta, p.C, p.crs, p.ccs, st);
    if (rc) return rc;
    if (p.K <= 0 || p.alpha == (T) 0) return 0;
    rc = launch_saso_binned<T>(p, st);          // register-resident output, binned entry lists (saso_binned.cu)
    if (rc >= 0) return rc;
    // only minor-axis vectors that intersect the window are visited
    const int64_t w0 = p.major_is_rows ? p.co_s : p.ro_s;
    const int64_t wn = p.major_is_rows ? p.cs : p.rs;
    int64_t warps = wn;
    int64_t grid = (warps + 7) / 8;
    int64_t cap = (int64_t) sm_count() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    saso_apply_kernel<T><<<(unsigned) grid, 256, 0, st>>>(p, w0, w0 + wn);
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}
template int launch_saso_apply<float>(const SasoProblem<float>&, cudaStream_t);
template int launch_saso_apply<double>(const SasoProblem<double>&, cudaStream_t);

template <typename T>
int launch_coo_apply(const CooProblem<T>& p, cudaStream_t st) {
    if (p.P <= 0 || p.Q <= 0) return 0;
    int rc = launch_scale<T>(p.P, p.Q, p.beta, p.C, p.crs, p.ccs, st);
    if (rc) return rc;
    if (p.K <= 0 || p.nnz <= 0 || p.alpha == (T) 0) return 0;
    int64_t grid = (p.nnz + 7) / 8;
    int64_t cap = (int64_t) sm_count() * 8;
    if (grid > cap) grid = cap;
    if (p.idx_bytes == 4) coo_apply_kernel<T, int32_t><<<(unsigned) grid, 256, 0, st>>>(p);
    else coo_apply_kernel<T, int64_t><<<(unsigned) grid, 256, 0, st>>>(p);
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}
// CSR rowptr / CSC colptr -> one index per stored entry (warp per major index), so that compressed formats can use
// the COO kernel: out[e] = r for ptr[r] <= e < ptr[r + 1]
template <typename IDX>
__global__ void __launch_bounds__(256) expand_ptr_kernel(int64_t n_major, const IDX* __restrict__ ptr, IDX* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_major; r += nwarps)
        for (int64_t e = (int64_t) ptr[r] + lane; e < (int64_t) ptr[r + 1]; e += 32) out[e] = (IDX) r;
}

int launch_expand_ptr(int64_t n_major, const void* ptr, void* out, int idx_bytes, cudaStream_t st) {
    if (n_major <= 0) return 0;
    int64_t grid = (n_major + 7) / 8;
    const int64_t cap = (int64_t) sm_count() * 8;
    if (grid > cap) grid = cap;
    if (idx_bytes == 4) expand_ptr_kernel<int32_t><<<(unsigned) grid, 256, 0, st>>>(n_major, (const int32_t*) ptr, (int32_t*) out);
    else expand_ptr_kernel<int64_t><<<(unsigned) grid, 256, 0, st>>>(n_major, (const int64_t*) ptr, (int64_t*) out);
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}

template int launch_coo_apply<float>(const CooProblem<float>&, cudaStream_t);
template int launch_coo_apply<double>(const CooProblem<double>&, cudaStream_t);

}  // namespace rb
