// K3c: long-axis-sparse operators (LASO), the Axis::Long branch of fill_sparse_unpacked_nosub
// (RandBLAS/sparse_skops.hh:534-564) with its helpers sample_indices_iid_uniform (RandBLAS/util.hh:515-547) and
// laso_merge_long_axis_vector_coo_data (sparse_skops.hh:453-491).
//
// Reference semantics restated. The operator has dim_minor "long-axis vectors" (one per index of the short axis).
// Vector i draws vec_nnz (index, sign) pairs: pair j takes two words of Philox block seed + i*ceil(vec_nnz/2) + j/2
// (words 0,1 for even j, words 2,3 for odd j):
//     index = (sint_t) (dim_major * ((double) uneg11<float>(w_a) + 1.0) / 2.0)        util.hh:532-533
//     sign  = uneg11<float>(w_b) >= 0 ? +1 : -1   (<=> (int32) w_b >= 0)              util.hh:536
// Repeated indices inside a vector are merged into one entry with value sqrt(count) * (sign of the first
// occurrence) (sparse_skops.hh:484-488), so vectors have between 1 and vec_nnz entries and the COO arrays are
// compacted: nnz <= vec_nnz * dim_minor. The state advances by dim_minor * ceil(vec_nnz/2) blocks (:274-279).
//
// ORDER inside a vector: when a vector has no repeated index the reference leaves the entries in draw order; when it
// has, it rewrites them in std::unordered_map iteration order, which is a property of the C++ standard library in
// use, not of RandBLAS. This build emits FIRST-OCCURRENCE order in both cases. The (index, value) SET of every
// vector, the per-vector counts, nnz and the next state are identical to the reference's; the tests compare vectors
// with repeats as sets.
//
// Two passes (count, cub exclusive scan, write) because the output offset of vector i depends on all vectors
// before it; the draws are regenerated in the second pass (one Philox block per two entries: far cheaper than
// staging them). vec_nnz <= 32: one warp per vector, duplicates found with match.any. Larger vec_nnz: one thread
// per vector with its draws in a global scratch row.
#include <cub/device/device_scan.cuh>
#include "common.cuh"
#include "kernels.h"

namespace rb {

namespace {

__device__ __forceinline__ void laso_draw(const Ctr128& ctr, const PhiloxKey& key, int64_t vec, int64_t kblk, int j,
                                          double dN, int64_t& idx, int& negative) {
    const uint4 w = philox4x32_10(ctr_add(ctr, (uint64_t) (vec * kblk + (j >> 1))), key);
    const uint32_t wa = (j & 1) ? w.z : w.x, wb = (j & 1) ? w.w : w.y;
    const double u = __ddiv_rn(__dadd_rn((double) uneg11f(wa), 1.0), 2.0);     // uneg11_to_u01<double>, util.hh:510-512
    idx = (int64_t) __dmul_rn(dN, u);                                          // truncation toward zero
    negative = ((int32_t) wb < 0) ? 1 : 0;
}

template <typename VAL>
__device__ __forceinline__ VAL laso_value(int count, int negative) {
    VAL s;
    if constexpr (sizeof(VAL) == 4) s = __fsqrt_rn((float) count);
    else s = __dsqrt_rn((double) count);
    return negative ? -s : s;                                                  // sqrt(c) * loc2scale[ell]
}

// ---- vec_nnz <= 32: warp per vector ----
template <typename IDX, typename VAL, bool WRITE>
__global__ void __launch_bounds__(256) laso_warp_kernel(Ctr128 ctr, PhiloxKey key, int k, int64_t dim_major,
                                                        int64_t dim_minor, int64_t* __restrict__ counts,
                                                        const int64_t* __restrict__ offs, IDX* __restrict__ lng,
                                                        IDX* __restrict__ sht, VAL* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    const int64_t kblk = (k + 1) / 2;
    const double dN = (double) dim_major;
    const unsigned live = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
    for (int64_t v = warp; v < dim_minor; v += nwarps) {
        int64_t idx = 0;
        int neg = 0;
        if (lane < k) laso_draw(ctr, key, v, kblk, lane, dN, idx, neg);
        unsigned same = 0;
        if (lane < k) same = __match_any_sync(live, (unsigned long long) idx);
        const bool first = lane < k && (__ffs(same) - 1) == lane;
        const unsigned firsts = __ballot_sync(0xffffffffu, first);
        if constexpr (!WRITE) {
            if (lane == 0) counts[v] = __popc(firsts);
        } else {
            if (first) {
                const int64_t e = offs[v] + __popc(firsts & ((1u << lane) - 1u));
                lng[e] = (IDX) idx;
                sht[e] = (IDX) v;
                vals[e] = laso_value<VAL>(__popc(same), neg);
            }
        }
    }
}

// ---- any vec_nnz: thread per vector, draws kept in a global scratch row of length k ----
template <typename IDX, typename VAL, bool WRITE>
__global__ void laso_thread_kernel(Ctr128 ctr, PhiloxKey key, int64_t k, int64_t dim_major, int64_t dim_minor,
                                   int64_t* __restrict__ counts, const int64_t* __restrict__ offs, IDX* __restrict__ lng,
                                   IDX* __restrict__ sht, VAL* __restrict__ vals, int64_t* __restrict__ scratch) {
    const int64_t gtid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    int64_t* mine = scratch + gtid * k;
    const int64_t kblk = (k + 1) / 2;
    const double dN = (double) dim_major;
    for (int64_t v = gtid; v < dim_minor; v += (int64_t) gridDim.x * blockDim.x) {
        for (int64_t j = 0; j < k; ++j) {
            int64_t idx; int neg;
            laso_draw(ctr, key, v, kblk, (int) j, dN, idx, neg);
            mine[j] = (idx << 1) | neg;            // dim_major < 2^62
        }
        int64_t n = 0;
        for (int64_t j = 0; j < k; ++j) {
            const int64_t ij = mine[j] >> 1;
            bool first = true;
            for (int64_t t = 0; t < j && first; ++t) first = (mine[t] >> 1) != ij;
            if (!first) continue;
            if constexpr (WRITE) {
                int c = 1;
                for (int64_t t = j + 1; t < k; ++t) c += ((mine[t] >> 1) == ij) ? 1 : 0;
                const int64_t e = offs[v] + n;
                lng[e] = (IDX) ij;
                sht[e] = (IDX) v;
                vals[e] = laso_value<VAL>(c, (int) (mine[j] & 1));
            }
            ++n;
        }
        if constexpr (!WRITE) counts[v] = n;
    }
}

template <typename IDX, typename VAL>
int launch_laso_t(Ctr128 ctr, PhiloxKey key, int64_t k, int64_t dim_major, int64_t dim_minor, void* lng, void* sht,
                  void* vals, int64_t* nnz_host, cudaStream_t st) {
    int64_t* counts = (int64_t*) workspace(0, (size_t) (dim_minor + 1) * 8, st);
    int64_t* offs = (int64_t*) workspace(1, (size_t) (dim_minor + 1) * 8, st);
    if (!counts || !offs) return fail_cuda(cudaErrorMemoryAllocation, "LASO count workspace");
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, offs, (int) (dim_minor + 1), st);
    void* tmp = workspace(2, tmp_bytes, st);
    if (!tmp) return fail_cuda(cudaErrorMemoryAllocation, "scan workspace");
    RB_CUDA(cudaMemsetAsync(counts + dim_minor, 0, 8, st));     // so that offs[dim_minor] is the total
    const int64_t cap = (int64_t) sm_count() * 8;
    if (k <= 32) {
        int64_t grid = (dim_minor + 7) / 8;
        if (grid > cap) grid = cap;
        laso_warp_kernel<IDX, VAL, false><<<(unsigned) grid, 256, 0, st>>>(ctr, key, (int) k, dim_major, dim_minor, counts,
                                                                         nullptr, nullptr, nullptr, nullptr);
        count_launch();
        RB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, offs, (int) (dim_minor + 1), st));
        count_launch();
        laso_warp_kernel<IDX, VAL, true><<<(unsigned) grid, 256, 0, st>>>(ctr, key, (int) k, dim_major, dim_minor, nullptr,
                                                                        offs, (IDX*) lng, (IDX*) sht, (VAL*) vals);
        count_launch();
    } else {
        int64_t grid = (dim_minor + 127) / 128;
        if (grid > cap) grid = cap;
        int64_t* scratch = (int64_t*) workspace(7, (size_t) (grid * 128 * k) * 8, st);
        if (!scratch) return fail_cuda(cudaErrorMemoryAllocation, "LASO scratch");
        laso_thread_kernel<IDX, VAL, false><<<(unsigned) grid, 128, 0, st>>>(ctr, key, k, dim_major, dim_minor, counts,
                                                                           nullptr, nullptr, nullptr, nullptr, scratch);
        count_launch();
        RB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, offs, (int) (dim_minor + 1), st));
        count_launch();
        laso_thread_kernel<IDX, VAL, true><<<(unsigned) grid, 128, 0, st>>>(ctr, key, k, dim_major, dim_minor, nullptr, offs,
                                                                          (IDX*) lng, (IDX*) sht, (VAL*) vals, scratch);
        count_launch();
    }
    RB_CUDA(cudaGetLastError());
    RB_CUDA(cudaMemcpyAsync(nnz_host, offs + dim_minor, 8, cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace

// idxs_long: index along the long (major) axis; idxs_short: the vector's own index. nnz_host receives the entry count.
int launch_laso(Ctr128 ctr, PhiloxKey key, int64_t vec_nnz, int64_t dim_major, int64_t dim_minor, void* idxs_long,
                void* idxs_short, int idx_bytes, void* vals, int val_bytes, int64_t* nnz_host, cudaStream_t st) {
    if (dim_minor <= 0 || vec_nnz <= 0) { *nnz_host = 0; return 0; }
    if (dim_minor >= 0x7fffffffLL) return fail("LASO: more than 2^31 - 1 long-axis vectors are not supported");
    if (idx_bytes == 4) {
        if (val_bytes == 4) return launch_laso_t<int32_t, float>(ctr, key, vec_nnz, dim_major, dim_minor, idxs_long, idxs_short, vals, nnz_host, st);
        return launch_laso_t<int32_t, double>(ctr, key, vec_nnz, dim_major, dim_minor, idxs_long, idxs_short, vals, nnz_host, st);
    }
    if (val_bytes == 4) return launch_laso_t<int64_t, float>(ctr, key, vec_nnz, dim_major, dim_minor, idxs_long, idxs_short, vals, nnz_host, st);
    return launch_laso_t<int64_t, double>(ctr, key, vec_nnz, dim_major, dim_minor, idxs_long, idxs_short, vals, nnz_host, st);
}

}  // namespace rb
