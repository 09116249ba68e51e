// K3b: SASO apply with register-resident output ("owner" kernel), the fast path of sparse::lskges / rskges
// (RandBLAS/skge.hh:465-492, 598-626) when the short axis of the operator indexes the rows of the result:
//     C(P x Q) += alpha * X(P x K) * Y(K x Q),   X = window of a SASO operator with vec_nnz entries per column.
// The reference sorts the operator's COO arrays into CSC and does one axpy of length Q per nonzero
// (sparse_data/coo_spmm_impl.hh:53-105, csc_spmm_impl.hh:99-209). Here:
//
//   1. saso_entries_kernel regenerates the operator as a compact stream: one 32-bit word per nonzero,
//      (row << 1 | negative), vec_nnz words per column of X, in column order (Fisher-Yates exactly as
//      sparse_skops.hh:72-102, one sub-warp group of lanes per column). This is 4 bytes per nonzero instead of
//      the 20 of the reference's (int64, int64, float) COO triplets, and it is the only form of S that exists.
//   2. saso_owner_kernel: a CTA owns a 1024 x 32 tile of C in REGISTERS for the whole kernel (every 8-lane group
//      owns 8 rows x 32 columns, 4 columns per lane) and walks over chunks of up to 768 rows of Y. Per chunk the
//      32 columns it needs of those rows are brought in by TMA (double buffered, 128-byte rows), the chunk's
//      entries are counting-sorted by target row in shared memory, and every group then adds the Y rows of its
//      bins into its accumulators: one conflict-free 128-byte shared-memory read per (entry, group), no atomics
//      and no read-modify-write of C anywhere in the loop. The CTAs that share a row range of Y run side by
//      side, so A is read from HBM once. At the end the tiles are added into C (one red per element and CTA).
//
// Roofline: HBM, bytes of Y (the data matrix A) read once; the inner loop itself is bound by shared-memory
// bandwidth / issue slots (every element of A is added into vec_nnz accumulators: vec_nnz * 4 B of
// shared-memory reads per 4 B of HBM).
#include "common.cuh"
#include "kernels.h"
#include "tma.cuh"

namespace rb {

namespace {

constexpr int OW_THREADS = 1024;
constexpr int OW_W = 32;                      // columns of C per CTA = one 128-byte row of Y per quarter-warp
constexpr int OW_RPG = 8;                     // rows of C per 8-lane group
constexpr int OW_PT = (OW_THREADS / 8) * OW_RPG;   // 1024 rows of C per CTA
constexpr int OW_KMAX = 768;                  // rows of Y per chunk
constexpr int OW_BOXR = 64;                   // rows per TMA box
constexpr int OW_ENT = 6144;                  // entries per chunk
constexpr int OW_EPT = OW_ENT / OW_THREADS;   // entries binned per thread and chunk
constexpr uint32_t OW_YBYTES = OW_KMAX * OW_W * 4;
constexpr uint32_t OW_OFF_SORTED = 2 * OW_YBYTES;
constexpr uint32_t OW_OFF_CNT = OW_OFF_SORTED + OW_ENT * 4;
constexpr uint32_t OW_OFF_WSUM = OW_OFF_CNT + (OW_PT + 4) * 4;
constexpr uint32_t OW_OFF_BAR = OW_OFF_WSUM + 32 * 4;
constexpr uint32_t OW_SMEM = OW_OFF_BAR + 16 + 128;   // + alignment slack
constexpr uint32_t ENTRY_NONE = 0xffffffffu;

// ---- pre-pass: the operator as a stream of (row << 1 | negative) words -------------------------------------
// G lanes per vector (power of two >= k). All arithmetic in 32 bits: dim_major < 2^31 on this path.
template <int G>
__global__ void __launch_bounds__(256) saso_entries_kernel(Ctr128 ctr, PhiloxKey key, int k, uint32_t dim_major,
                                                           int64_t vec_lo, int64_t nvec, int64_t m0, int64_t P,
                                                           uint32_t* __restrict__ entries) {
    constexpr int VPW = 32 / G;
    const int lane = threadIdx.x & 31, sub = lane & (G - 1);
    const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    for (int64_t wv = warp * VPW; wv < nvec; wv += nwarps * VPW) {
        const int64_t v = wv + lane / G;
        const bool live = v < nvec && sub < k;
        uint32_t piv = 0, w1 = 0;
        if (live) {
            const uint4 w = philox4x32_10(ctr_add(ctr, (uint64_t) ((vec_lo + v) * k + sub)), key);
            piv = (uint32_t) sub + w.x % (dim_major - (uint32_t) sub);      // sparse_skops.hh:78
            w1 = w.y;
        }
        // value at position piv after swaps 0..sub-1 of an identity permutation: walk the swaps backwards
        uint32_t pos = piv;
        for (int t = k - 2; t >= 0; --t) {
            const uint32_t pt = __shfl_sync(0xffffffffu, piv, t, G);
            if (t < sub) {
                if (pos == (uint32_t) t) pos = pt;
                else if (pos == pt) pos = (uint32_t) t;
            }
        }
        if (live) {
            const int64_t r = (int64_t) pos - m0;
            entries[v * k + sub] = (r >= 0 && r < P) ? (((uint32_t) r << 1) | (w1 & 1u)) : ENTRY_NONE;   // :84-88
        }
    }
}

struct OwnerArgs {
    const uint32_t* entries;
    int64_t nvec;        // columns of X (rows of Y) in this launch
    int64_t nchunks;
    int k;               // entries per column
    uint32_t kmagic;     // floor(2^32 / k) + 1: e / k == umulhi(e, kmagic) for e * k < 2^32; 0 when k == 1
    int Kc;              // rows of Y per chunk, a multiple of OW_BOXR, Kc * k <= OW_ENT
    int G;               // CTAs that share one tile of C (they split the chunks)
    int64_t P, Q;
    float alpha;
    float* C;
    int64_t crs;
    int c_vec4;          // rows of C are 16-byte aligned
};

__global__ void __launch_bounds__(OW_THREADS, 1) saso_owner_kernel(const __grid_constant__ CUtensorMap tmY,
                                                                   const OwnerArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = tma::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 127u) & ~127u) - raw);
    uint32_t* sorted = reinterpret_cast<uint32_t*>(smem + OW_OFF_SORTED);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + OW_OFF_CNT);
    uint32_t* wsum = reinterpret_cast<uint32_t*>(smem + OW_OFF_WSUM);
    const uint32_t bar0 = tma::smem_u32(smem + OW_OFF_BAR);
    const uint32_t ybuf0 = tma::smem_u32(smem);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gi = tid >> 3, l8 = tid & 7;
    const int col0 = (int) blockIdx.x * OW_W;
    const int64_t row0 = (int64_t) blockIdx.y * OW_PT;
    const uint32_t row0_u = (uint32_t) row0;
    const int g = (int) blockIdx.z;
    const int nbox = a.Kc / OW_BOXR;
    const uint32_t chunk_bytes = (uint32_t) a.Kc * OW_W * 4;
    const int64_t ent_per_chunk = (int64_t) a.Kc * a.k;
    const int64_t ent_total = a.nvec * a.k;

    if (tid == 0) {
        tma::mbar_init(bar0, 1);
        tma::mbar_init(bar0 + 8, 1);
        tma::mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int64_t c, int b) {
        const uint32_t bar = bar0 + 8u * (uint32_t) b;
        tma::mbar_arrive_expect_tx(bar, chunk_bytes);
        const uint32_t dst = ybuf0 + (uint32_t) b * OW_YBYTES;
        for (int x = 0; x < nbox; ++x)
            tma::load_2d(dst + (uint32_t) x * (OW_BOXR * OW_W * 4), &tmY, bar, col0, (int) (c * a.Kc + x * OW_BOXR));
    };
    auto fetch = [&](int64_t c, uint32_t (&ent)[OW_EPT]) {
        const int64_t e0 = c * ent_per_chunk;
#pragma unroll
        for (int i = 0; i < OW_EPT; ++i) {
            const int e = tid + i * OW_THREADS;
            ent[i] = (c < a.nchunks && e < ent_per_chunk && e0 + e < ent_total) ? __ldg(a.entries + e0 + e) : ENTRY_NONE;
        }
    };

    float acc[OW_RPG][4];
#pragma unroll
    for (int q = 0; q < OW_RPG; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;

    uint32_t nxt[OW_EPT];
    fetch(g, nxt);
    if (tid == 0 && g < a.nchunks) issue(g, 0);

    int it = 0;
    for (int64_t c = g; c < a.nchunks; c += a.G, ++it) {
        const int b = it & 1;
        if (tid == 0 && c + a.G < a.nchunks) issue(c + a.G, b ^ 1);   // that buffer was released by the last barrier

        // ---- counting sort of this chunk's entries by target row (only rows of this CTA's tile) ----
        cnt[tid] = 0;
        __syncthreads();
        uint32_t pos[OW_EPT];
#pragma unroll
        for (int i = 0; i < OW_EPT; ++i) {
            const uint32_t r = (nxt[i] >> 1) - row0_u;          // ENTRY_NONE >> 1 is beyond every tile (P < 2^30)
            pos[i] = (r < (uint32_t) OW_PT) ? atomicAdd(&cnt[r], 1u) : 0u;
        }
        __syncthreads();
        {
            const uint32_t v = cnt[tid];
            uint32_t incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const uint32_t s = wsum[lane];
                uint32_t si = s;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
                    if (lane >= o) si += t;
                }
                wsum[lane] = si - s;
            }
            __syncthreads();
            const uint32_t off = incl - v + wsum[warp];
            cnt[tid] = off;
            if (tid == OW_THREADS - 1) cnt[OW_PT] = off + v;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < OW_EPT; ++i) {
            const uint32_t r = (nxt[i] >> 1) - row0_u;
            if (r < (uint32_t) OW_PT) {
                const uint32_t e = (uint32_t) (tid + i * OW_THREADS);
                const uint32_t vloc = a.kmagic ? __umulhi(e, a.kmagic) : e;  // column of X within the chunk
                sorted[cnt[r] + pos[i]] = (vloc * (uint32_t) (OW_W * 4)) | ((nxt[i] & 1u) << 31);
            }
        }
        fetch(c + a.G, nxt);          // next chunk's entries travel while this chunk is accumulated
        __syncthreads();

        // ---- accumulate: group gi owns rows gi*8 .. gi*8+7; lane l8 owns 4 of the 32 columns ----
        tma::mbar_wait(bar0 + 8u * (uint32_t) b, (uint32_t) ((it >> 1) & 1));
        const uint8_t* ybase = smem + (uint32_t) b * OW_YBYTES + l8 * 16;
        uint32_t o[OW_RPG + 1];
#pragma unroll
        for (int q = 0; q <= OW_RPG; ++q) o[q] = cnt[gi * OW_RPG + q];
#pragma unroll
        for (int q = 0; q < OW_RPG; ++q) {
#pragma unroll 2
            for (uint32_t e = o[q]; e < o[q + 1]; ++e) {
                const uint32_t p = sorted[e];
                const float4 y = *reinterpret_cast<const float4*>(ybase + (p & 0x7fffffffu));
                const float s = __uint_as_float((p & 0x80000000u) | 0x3f800000u);
                acc[q][0] = fmaf(y.x, s, acc[q][0]);
                acc[q][1] = fmaf(y.y, s, acc[q][1]);
                acc[q][2] = fmaf(y.z, s, acc[q][2]);
                acc[q][3] = fmaf(y.w, s, acc[q][3]);
            }
        }
        __syncthreads();
    }

    // ---- add the tile into C (beta was applied beforehand; G CTAs share the tile) ----
    const int64_t col = (int64_t) col0 + l8 * 4;
#pragma unroll
    for (int q = 0; q < OW_RPG; ++q) {
        const int64_t row = row0 + (int64_t) gi * OW_RPG + q;
        if (row >= a.P || col >= a.Q) continue;
        float* cp = a.C + row * a.crs + col;
        const float v0 = a.alpha * acc[q][0], v1 = a.alpha * acc[q][1], v2 = a.alpha * acc[q][2], v3 = a.alpha * acc[q][3];
        if (a.c_vec4 && col + 3 < a.Q) {
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(cp), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
        } else {
            atomicAdd(cp, v0);
            if (col + 1 < a.Q) atomicAdd(cp + 1, v1);
            if (col + 2 < a.Q) atomicAdd(cp + 2, v2);
            if (col + 3 < a.Q) atomicAdd(cp + 3, v3);
        }
    }
}

template <int G>
void launch_entries(const SasoProblem<float>& p, int64_t vec_lo, int64_t nvec, int64_t m0, uint32_t* entries,
                    cudaStream_t st) {
    const int64_t vpb = 8 * (32 / G);       // vectors per 256-thread CTA and iteration
    int64_t grid = (nvec + vpb - 1) / vpb;
    const int64_t cap = (int64_t) sm_count() * 16;
    if (grid > cap) grid = cap;
    saso_entries_kernel<G><<<(unsigned) grid, 256, 0, st>>>(p.ctr, p.key, (int) p.vec_nnz, (uint32_t) p.dim_major, vec_lo,
                                                           nvec, m0, p.P, entries);
}

}  // namespace

// Returns 0 if the product was computed, -1 if this path does not take the problem (caller falls back), >0 on error.
// C must have been beta-scaled by the caller.
int launch_saso_owner_f32(const SasoProblem<float>& p, cudaStream_t st) {
    if (get_option("saso_path") == 1) return -1;                       // 1 = force the atomic kernel
    const bool scatter = p.major_is_rows ? !p.x_is_transposed : p.x_is_transposed;   // short axis <-> rows of C
    if (!scatter) return -1;
    if (p.ycs != 1 || p.ccs != 1) return -1;
    if ((reinterpret_cast<uintptr_t>(p.Y) & 15) != 0 || (p.yrs & 3) != 0) return -1;   // TMA alignment rules
    if (p.vec_nnz > 32 || p.dim_major >= 0x7fffffffLL || p.P >= 0x40000000LL) return -1;
    if (p.Q > 0x7fffffffLL || p.yrs > 0x3fffffffLL) return -1;
    const int64_t w0 = p.major_is_rows ? p.co_s : p.ro_s;      // first column of X in operator coordinates
    const int64_t m0 = p.major_is_rows ? p.ro_s : p.co_s;      // first row of X
    const int64_t nvec = p.K;
    if (get_option("saso_path") < 2 && nvec * p.vec_nnz < 32768) return -1;   // small: one launch of the atomic kernel wins
    tma::EncodeTiledFn enc = tma::encode_tiled_fn();
    if (!enc) return -1;
    {
        static DevOnce attr_done;
        if (attr_done.need()) {
            if (cudaFuncSetAttribute(saso_owner_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OW_SMEM) != cudaSuccess) {
                cudaGetLastError();
                return -1;
            }
            attr_done.done();
        }
    }
    const int k = (int) p.vec_nnz;
    int Kc = (OW_ENT / k) / OW_BOXR * OW_BOXR;
    if (Kc > OW_KMAX) Kc = OW_KMAX;
    const int64_t ns = (p.Q + OW_W - 1) / OW_W, np = (p.P + OW_PT - 1) / OW_PT;
    if (ns > 0x7fffffffLL || np > 65535) return -1;
    const int sms = sm_count();
    // segments bound the entry workspace (4 B per nonzero) to 1 GiB
    int64_t seg_vecs = ((int64_t) 1 << 28) / k / Kc * Kc;
    for (int64_t v0 = 0; v0 < nvec; v0 += seg_vecs) {
        const int64_t nv = (nvec - v0 < seg_vecs) ? nvec - v0 : seg_vecs;
        uint32_t* entries = (uint32_t*) workspace(7, (size_t) nv * k * 4, st);
        if (!entries) return fail_cuda(cudaErrorMemoryAllocation, "SASO entry workspace");
        if (k <= 1) launch_entries<1>(p, w0 + v0, nv, m0, entries, st);
        else if (k <= 2) launch_entries<2>(p, w0 + v0, nv, m0, entries, st);
        else if (k <= 4) launch_entries<4>(p, w0 + v0, nv, m0, entries, st);
        else if (k <= 8) launch_entries<8>(p, w0 + v0, nv, m0, entries, st);
        else if (k <= 16) launch_entries<16>(p, w0 + v0, nv, m0, entries, st);
        else launch_entries<32>(p, w0 + v0, nv, m0, entries, st);
        count_launch();
        RB_CUDA(cudaGetLastError());

        CUtensorMap tm;
        const cuuint64_t gdim[2] = {(cuuint64_t) p.Q, (cuuint64_t) nv};
        const cuuint64_t gstr[1] = {(cuuint64_t) p.yrs * 4ull};
        const cuuint32_t box[2] = {OW_W, OW_BOXR};
        const cuuint32_t estr[2] = {1, 1};
        CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.Y + v0 * p.yrs), gdim, gstr, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return v0 == 0 ? -1 : fail("cuTensorMapEncodeTiled failed for the SASO apply");

        OwnerArgs a;
        a.entries = entries;
        a.nvec = nv;
        a.k = k;
        a.kmagic = (k == 1) ? 0u : (uint32_t) (0x100000000ull / (uint64_t) k) + 1u;
        a.Kc = Kc;
        a.nchunks = (nv + Kc - 1) / Kc;
        int64_t G = sms / (ns * np);
        if (G < 1) G = 1;
        if (G > a.nchunks) G = a.nchunks;
        if (G > 65535) G = 65535;
        a.G = (int) G;
        a.P = p.P; a.Q = p.Q;
        a.alpha = p.alpha;
        a.C = p.C; a.crs = p.crs;
        a.c_vec4 = ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && (p.crs & 3) == 0) ? 1 : 0;
        dim3 grid((unsigned) ns, (unsigned) np, (unsigned) G);
        saso_owner_kernel<<<grid, OW_THREADS, OW_SMEM, st>>>(tm, a);
        count_launch();
        count_owner_launch();
        RB_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace rb
