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
#include <type_traits>

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

// ---- thread per vector (K = 2, 4, 8, 16 entries, dim_major < 2^31) -------------------------------------------------
// The lane-per-entry kernel above spends most of its time on the 16-lane integer pipe: per nonzero a masked backward
// trace through 7 shuffles (3 integer instructions per step), a 128-bit counter addition and 64-bit index arithmetic
// (~150 instructions per warp-level nonzero, ALU pipe 73% busy). With one THREAD per vector the K pivots live in
// registers: the trace of entry j is j compares-and-selects with compile-time bounds (K (K - 1) / 2 per vector, no
// shuffles, no masks), the counter of entry j is the vector's counter plus j (one 128-bit addition per vector), the K
// divisors' reciprocals are kernel parameters (constant-bank operands), and a thread's K consecutive entries of each COO
// array leave as 32-byte (STG.256) or 16-byte stores: every store instruction writes whole sectors.
struct SasoVecMagic {
    uint32_t m[16];      // fastmod_magic(dim_major - j)
};

__device__ __forceinline__ void st256(void* p, unsigned long long a, unsigned long long b, unsigned long long c,
                                      unsigned long long d) {
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ void st256(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                      uint32_t g, uint32_t h) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f),
                 "r"(g), "r"(h)
                 : "memory");
}
// the bits of an index or value as the unsigned integer of its size
template <typename T>
__device__ __forceinline__ auto bits_of(T x) {
    if constexpr (sizeof(T) == 8) {
        if constexpr (std::is_floating_point<T>::value) return (unsigned long long) __double_as_longlong((double) x);
        else return (unsigned long long) x;
    } else {
        if constexpr (std::is_floating_point<T>::value) return (uint32_t) __float_as_uint((float) x);
        else return (uint32_t) x;
    }
}

// N consecutive elements of T from registers to dst (dst is 32-byte aligned whenever N * sizeof(T) is a multiple of 32,
// 16-byte aligned when it is a multiple of 16: the launcher checks the base pointers)
template <typename T, int N>
__device__ __forceinline__ void store_run(T* dst, const T (&x)[N]) {
    constexpr int BYTES = N * (int) sizeof(T);
    if constexpr (BYTES % 32 == 0 && sizeof(T) == 8) {
#pragma unroll
        for (int c = 0; c < N / 4; ++c)
            st256(dst + 4 * c, bits_of(x[4 * c]), bits_of(x[4 * c + 1]), bits_of(x[4 * c + 2]), bits_of(x[4 * c + 3]));
    } else if constexpr (BYTES % 32 == 0) {
#pragma unroll
        for (int c = 0; c < N / 8; ++c)
            st256(dst + 8 * c, bits_of(x[8 * c]), bits_of(x[8 * c + 1]), bits_of(x[8 * c + 2]), bits_of(x[8 * c + 3]),
                  bits_of(x[8 * c + 4]), bits_of(x[8 * c + 5]), bits_of(x[8 * c + 6]), bits_of(x[8 * c + 7]));
    } else if constexpr (BYTES % 16 == 0 && sizeof(T) == 8) {
#pragma unroll
        for (int c = 0; c < N / 2; ++c)
            *reinterpret_cast<ulonglong2*>(dst + 2 * c) = make_ulonglong2(bits_of(x[2 * c]), bits_of(x[2 * c + 1]));
    } else if constexpr (BYTES % 16 == 0) {
#pragma unroll
        for (int c = 0; c < N / 4; ++c)
            *reinterpret_cast<uint4*>(dst + 4 * c) = make_uint4(bits_of(x[4 * c]), bits_of(x[4 * c + 1]), bits_of(x[4 * c + 2]),
                                                                bits_of(x[4 * c + 3]));
    } else {
#pragma unroll
        for (int c = 0; c < N; ++c) dst[c] = x[c];
    }
}

template <typename IDX, typename VAL, int K>
__global__ void __launch_bounds__(256) saso_fill_vec_kernel(Ctr128 ctr, PhiloxKey key, uint32_t dim_major, int64_t dim_minor,
                                                            const __grid_constant__ SasoVecMagic mg, IDX* __restrict__ maj,
                                                            IDX* __restrict__ mnr, VAL* __restrict__ vals) {
    const int64_t nthreads = (int64_t) gridDim.x * blockDim.x;
    for (int64_t v = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; v < dim_minor; v += nthreads) {
        const Ctr128 c0 = ctr_add(ctr, (uint64_t) v * (uint64_t) K);
        const bool nocarry = c0.c0 <= 0xffffffffu - (uint32_t) K;      // entry j's counter differs in the low word only
        uint32_t piv[K], neg[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            Ctr128 cj = c0;
            if (nocarry) cj.c0 = c0.c0 + (uint32_t) j;
            else cj = ctr_add(c0, (uint64_t) j);
            const uint4 w = philox4x32_10(cj, key);
            piv[j] = (uint32_t) j + fastmod(w.x, dim_major - (uint32_t) j, mg.m[j]);          // sparse_skops.hh:78
            neg[j] = w.y & 1u;
        }
        // value at position piv[j] after swaps 0..j-1 of an identity permutation (see saso_vector_warp)
        IDX om[K], on[K];
        VAL ov[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            uint32_t pos = piv[j];
#pragma unroll
            for (int t = j - 1; t >= 0; --t)
                if (pos == piv[t]) pos = (uint32_t) t;
            om[j] = (IDX) pos;
            on[j] = (IDX) v;
            ov[j] = neg[j] ? (VAL) -1 : (VAL) 1;
        }
        store_run<IDX, K>(maj + v * K, om);
        if (mnr) store_run<IDX, K>(mnr + v * K, on);
        if (vals) store_run<VAL, K>(vals + v * K, ov);
    }
}

template <typename IDX, typename VAL, int K>
static void launch_saso_vec(Ctr128 ctr, PhiloxKey key, int64_t dim_major, int64_t dim_minor, void* maj, void* mnr, void* vals,
                            cudaStream_t st) {
    SasoVecMagic mg;
    for (int j = 0; j < 16; ++j) mg.m[j] = (j < K && dim_major - j > 0) ? 0xffffffffu / (uint32_t) (dim_major - j) : 0u;
    int64_t grid = (dim_minor + 255) / 256;
    const int64_t cap = (int64_t) sm_count() * 32;
    if (grid > cap) grid = cap;
    saso_fill_vec_kernel<IDX, VAL, K><<<(unsigned) grid, 256, 0, st>>>(ctr, key, (uint32_t) dim_major, dim_minor, mg, (IDX*) maj,
                                                                       (IDX*) mnr, (VAL*) vals);
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
    // saso_fill_path: 0 auto (thread per vector for k = 2, 4, 8, 16; lane per entry otherwise), 1 warp per vector,
    // 2 lane per entry always
    const bool aligned32 = ((reinterpret_cast<uintptr_t>(maj) | reinterpret_cast<uintptr_t>(mnr) | reinterpret_cast<uintptr_t>(vals)) & 31) == 0;
    if ((k == 2 || k == 4 || k == 8 || k == 16) && dim_major < 0x7fffffffLL && dim_major >= k && aligned32 &&
        get_option("saso_fill_path") == 0) {
        switch (k) {
            case 2: launch_saso_vec<IDX, VAL, 2>(ctr, key, dim_major, dim_minor, maj, mnr, vals, st); break;
            case 4: launch_saso_vec<IDX, VAL, 4>(ctr, key, dim_major, dim_minor, maj, mnr, vals, st); break;
            case 8: launch_saso_vec<IDX, VAL, 8>(ctr, key, dim_major, dim_minor, maj, mnr, vals, st); break;
            default: launch_saso_vec<IDX, VAL, 16>(ctr, key, dim_major, dim_minor, maj, mnr, vals, st); break;
        }
    } else if (k <= 32 && dim_major < 0x7fffffffLL && (get_option("saso_fill_path") & ~2) == 0) {
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
