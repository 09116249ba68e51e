// K3b (second generation): SASO apply with a register-resident output tile and a ONE-TIME binning pre-pass.
// Fast path of sparse::lskges / rskges (RandBLAS/skge.hh:465-492, 598-626) when the short axis of the operator
// indexes the rows of the result:
//     C(P x Q) += alpha * X(P x K) * Y(K x Q),   X = window of a SASO operator with vec_nnz entries per column.
// The reference deep-copies the operator's COO arrays, std::sorts them into CSC and does one axpy of length Q per
// nonzero (sparse_data/coo_spmm_impl.hh:53-105, csc_spmm_impl.hh:99-209). Here:
//
//   1. saso_bin_kernel (one CTA per chunk of Kc columns of X = Kc rows of Y): regenerates the chunk's nonzeros
//      (Fisher-Yates exactly as sparse_skops.hh:72-102, one sub-warp group of lanes per column), counting-sorts
//      them by target row in shared memory and writes, per chunk, the sorted list (one 32-bit word per nonzero:
//      byte offset of the Y row inside the chunk's shared-memory image | sign << 31) and the row offsets (u16).
//      This is the only form of S that ever exists (4 B + ~6/vec_nnz B per nonzero instead of the 20 B of the
//      reference's COO triplets), and it is built once (a first design re-sorted every chunk in each of the
//      ns x np CTAs that needed it: 9.9 ms at the C4 shape instead of 5.5).
//   2. saso_binned_kernel: a CTA owns a 1024 x 32 tile of C in REGISTERS for the whole kernel and walks over its share of
//      the chunks. Default (RL, "saso_rows" = 1): thread t owns row t of the tile, all 32 columns, and walks that row's
//      list -- a warp walks 32 lists in lockstep, 8 conflict-free LDS.128 and 32 FMAs per lane and step. First form
//      ("saso_rows" = 0): every 8-lane group owns 8 rows x 32 columns, 4 columns per lane, one data-dependent loop per row. A two-stage
//      TMA pipeline brings in, per chunk, the 32 columns it needs of the Kc rows of Y (2D tensor map, 128-byte
//      rows), the part of the sorted list that targets its rows and the matching offsets (bulk copies). Warps
//      run decoupled: full[] mbarriers signal arrival, the last warp to finish a stage (elected through a
//      shared-memory counter, ordered by an empty[] mbarrier) re-arms it with the next chunk. No __syncthreads
//      in the loop, no atomics on C, conflict-free 16-byte shared-memory reads (RL: lane l reads chunk j ^ (l & 7) of
//      its entry's row in sub-step j; groups: one 128-byte row per (entry, group)).
//
// Roofline: HBM, bytes of Y (the data matrix A) read once. The inner loop itself is bound by shared-memory read
// bandwidth and issue slots: every element of A is added into vec_nnz accumulators, i.e. vec_nnz * 4 B of
// shared-memory reads per 4 B of HBM.
#include "common.cuh"
#include "kernels.h"
#include "tma.cuh"

namespace rb {

namespace {

constexpr int BN_THREADS = 1024;
constexpr int BN_W = 32;                        // 4-byte words per row slice: a CTA owns 128 bytes of every row of C, i.e.
                                                // 32 float or 16 double columns (one 128-byte row of Y per 8-lane group)
constexpr int BN_RPG = 8;                       // rows of C per 8-lane group
constexpr int BN_PT = (BN_THREADS / 8) * BN_RPG;   // 1024 rows of C per CTA
constexpr int BN_KMAX = 704;                    // rows of Y per chunk
constexpr int BN_BOXR = 64;                     // rows per TMA box
constexpr int BN_ENT = 5632;                    // nonzeros per chunk (Kc * vec_nnz <= BN_ENT)
constexpr int BN_ENT_CAP = BN_ENT + 8;          // per-chunk stride of the sorted list in global memory (entries)
constexpr int BN_PMAX = 8192;                   // rows of C the binning kernel can count in shared memory
constexpr uint32_t BN_YBYTES = BN_KMAX * BN_W * 4;          // 90112
constexpr uint32_t BN_LBYTES = BN_ENT * 4 + 16;             // list + slack for the 16-byte aligned start
constexpr uint32_t BN_OCOUNT = BN_PT + 8;                   // u16 offsets copied per chunk and row tile (1025 used)
constexpr uint32_t BN_OBYTES = BN_OCOUNT * 2;               // 2064
constexpr uint32_t BN_OFF_LIST = BN_YBYTES;
constexpr uint32_t BN_OFF_OFFS = BN_OFF_LIST + BN_LBYTES;
constexpr uint32_t BN_STAGE = ((BN_OFF_OFFS + BN_OBYTES + 127u) / 128u) * 128u;   // 114816
constexpr uint32_t BN_OFF_BAR = 2 * BN_STAGE;               // full[2], empty[2], cnt[2]
constexpr uint32_t BN_SMEM = BN_OFF_BAR + 64 + 128;         // + alignment slack

constexpr int BIN_THREADS = 512;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// One list word: (byte offset of the Y row in the chunk image) / 2 | negative << 31. The address is base + (p << 1)
// (the shift drops the sign bit), the multiplier +-1.0f is (p & 0x80000000) | 0x3f800000: one LOP3.
__device__ __forceinline__ float word_sign(uint32_t p) {
    uint32_t s;
    asm("lop3.b32 %0, %1, 0x80000000, 0x3f800000, 0xEA;" : "=r"(s) : "r"(p));
    return __uint_as_float(s);
}

struct BinArgs {
    Ctr128 ctr;
    PhiloxKey key;
    int k;                 // entries per column of X
    uint32_t dim_major;
    int64_t vec_lo;        // first column of X in operator coordinates
    int64_t nvec;          // columns of X in this launch
    int64_t m0;            // first row of X in operator coordinates
    int64_t P;             // rows of X
    int Kc;                // columns per chunk
    int64_t nchunks;
    int Prows;             // np * BN_PT
    int Ppad;              // Prows + 8: per-chunk stride of offs16
    uint32_t* sorted;      // [nchunks][BN_ENT_CAP]
    uint16_t* offs16;      // [nchunks][Ppad]
    uint32_t magic[16];    // VEC kernels: fastmod_magic(dim_major - j), constant-bank operands
};

// G lanes per column of X (power of two >= k). All index arithmetic in 32 bits: dim_major < 2^31 on this path.
// VEC (k == G = 2, 4, 8, 16): one THREAD regenerates a whole column -- the pivots stay in its registers, the backward trace
// has compile-time bounds (no shuffles, no masks) and a column needs one 128-bit counter addition (the generation scheme of
// saso_fill_vec_kernel, sparse_ops.cu; the lane-per-entry form spent ~280 instructions per entry, most of them on the
// 16-lane integer pipe).
template <int G, bool VEC = false, int NT = BIN_THREADS, int MINB = 1>
__global__ void __launch_bounds__(NT, MINB) saso_bin_kernel(const __grid_constant__ BinArgs a) {
    extern __shared__ __align__(16) uint32_t bsm[];
    uint32_t* cnt = bsm;                               // [Prows + 1] counters, then running offsets
    uint32_t* raw = cnt + a.Prows + 32;                // [BN_ENT] (row << 1 | negative), ~0 = outside the window
    uint32_t* srt = raw + BN_ENT;                      // [BN_ENT] sorted words
    constexpr int BIN_THREADS = NT;                    // (shadows the default thread count in this kernel)
    __shared__ uint32_t wsum[BIN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = lane & (G - 1);
    constexpr int VPP = BIN_THREADS / G;               // columns per pass
    const int per_thr = (a.Prows + BIN_THREADS - 1) / BIN_THREADS;     // counters scanned per thread
    const uint32_t dv = sub < a.k ? a.dim_major - (uint32_t) sub : 1u, dm = fastmod_magic(dv);

    for (int64_t c = blockIdx.x; c < a.nchunks; c += gridDim.x) {
        const int64_t v0 = c * a.Kc;
        const int nv = (int) min((int64_t) a.Kc, a.nvec - v0);
        for (int i = tid; i < a.Prows; i += BIN_THREADS) cnt[i] = 0;
        __syncthreads();
        // ---- generate the chunk's nonzeros, count them per target row ----
        if constexpr (VEC) {
            for (int vl = tid; vl < nv; vl += BIN_THREADS) {
                const Ctr128 c0 = ctr_add(a.ctr, (uint64_t) (a.vec_lo + v0 + vl) * (uint64_t) G);
                const bool nocarry = c0.c0 <= 0xffffffffu - (uint32_t) G;
                uint32_t piv[G], neg[G];
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    Ctr128 cj = c0;
                    if (nocarry) cj.c0 = c0.c0 + (uint32_t) j;
                    else cj = ctr_add(c0, (uint64_t) j);
                    const uint4 w = philox4x32_10(cj, a.key);
                    piv[j] = (uint32_t) j + fastmod(w.x, a.dim_major - (uint32_t) j, a.magic[j]);   // sparse_skops.hh:78
                    neg[j] = w.y & 1u;
                }
                uint32_t out[G];
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    uint32_t pos = piv[j];
#pragma unroll
                    for (int t = j - 1; t >= 0; --t)
                        if (pos == piv[t]) pos = (uint32_t) t;
                    const int64_t r = (int64_t) pos - a.m0;
                    const bool in = r >= 0 && r < a.P;
                    out[j] = in ? (((uint32_t) r << 1) | neg[j]) : 0xffffffffu;                       // :84-88
                    if (in) atomicAdd(&cnt[r], 1u);
                }
                if constexpr (G % 4 == 0) {
#pragma unroll
                    for (int q = 0; q < G / 4; ++q)
                        *reinterpret_cast<uint4*>(raw + vl * G + 4 * q) = make_uint4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
                } else {
                    *reinterpret_cast<uint2*>(raw + vl * G) = make_uint2(out[0], out[1]);
                }
            }
        } else
        for (int vb = 0; vb < nv; vb += VPP) {
            const int vl = vb + tid / G;
            const bool live = vl < nv && sub < a.k;
            uint32_t piv = 0, w1 = 0;
            if (live) {
                const uint4 w = philox4x32_10(ctr_add(a.ctr, (uint64_t) ((a.vec_lo + v0 + vl) * a.k + sub)), a.key);
                piv = (uint32_t) sub + fastmod(w.x, dv, dm);                      // sparse_skops.hh:78
                w1 = w.y;
            }
            // value at position piv after swaps 0..sub-1 of an identity permutation: walk the swaps backwards
            uint32_t pos = piv;
#pragma unroll
            for (int t = G - 2; t >= 0; --t) {
                const uint32_t pt = __shfl_sync(0xffffffffu, piv, t, G);
                if (t < sub && pos == pt) pos = (uint32_t) t;
            }
            if (live) {
                const int64_t r = (int64_t) pos - a.m0;
                const bool in = r >= 0 && r < a.P;
                raw[vl * a.k + sub] = in ? (((uint32_t) r << 1) | (w1 & 1u)) : 0xffffffffu;   // :84-88
                if (in) atomicAdd(&cnt[r], 1u);
            }
        }
        __syncthreads();
        // ---- exclusive scan of the counters (per_thr consecutive counters per thread) ----
        {
            uint32_t loc = 0;
            for (int i = 0; i < per_thr; ++i)
                if (tid * per_thr + i < a.Prows) loc += cnt[tid * per_thr + i];
            uint32_t incl = loc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const uint32_t s = (lane < BIN_THREADS / 32) ? wsum[lane] : 0u;
                uint32_t si = s;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
                    if (lane >= o) si += t;
                }
                if (lane < BIN_THREADS / 32) wsum[lane] = si - s;
            }
            __syncthreads();
            uint32_t run = incl - loc + wsum[warp];
            for (int i = 0; i < per_thr; ++i) {
                if (tid * per_thr + i >= a.Prows) break;
                const uint32_t v = cnt[tid * per_thr + i];
                cnt[tid * per_thr + i] = run;
                run += v;
            }
        }
        __syncthreads();
        // ---- scatter: after this pass cnt[r] is the END of row r's list ----
        const int ne = nv * a.k;
        for (int e = tid; e < ne; e += BIN_THREADS) {
            const uint32_t w = raw[e];
            if (w != 0xffffffffu) {
                const uint32_t dst = atomicAdd(&cnt[w >> 1], 1u);
                const uint32_t vl = VEC ? (uint32_t) e / (uint32_t) G : (uint32_t) e / (uint32_t) a.k;   // VEC: k == G, a shift
                srt[dst] = (vl * (uint32_t) (BN_W * 2)) | ((w & 1u) << 31);
            }
        }
        __syncthreads();
        // ---- write out: sorted words (coalesced) and u16 row offsets (offs[0] = 0, offs[r + 1] = end of row r) ----
        const uint32_t total = cnt[a.Prows - 1];
        uint32_t* gs = a.sorted + c * BN_ENT_CAP;
        for (uint32_t e = tid; e < ((total + 3u) & ~3u); e += BIN_THREADS) gs[e] = (e < total) ? srt[e] : 0u;
        uint16_t* go = a.offs16 + c * a.Ppad;
        for (int i = tid; i < a.Ppad; i += BIN_THREADS)
            go[i] = (uint16_t) (i == 0 ? 0u : (i <= a.Prows ? cnt[i - 1] : total));
        __syncthreads();
    }
}

template <typename T>
struct BinnedArgs {
    const uint32_t* sorted;
    const uint16_t* offs16;
    int nchunks;
    int Kc;              // rows of Y per chunk, a multiple of BN_BOXR
    int Ppad;
    int G;               // CTAs that share one tile of C (they split the chunks)
    int64_t P, Q;
    T alpha;
    T* C;
    int64_t crs;
    int c_vec4;          // rows of C are 16-byte aligned
};

// 16 bytes of a Y row and of the accumulators of one row of C: 4 floats or 2 doubles per lane
template <typename T> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int N = 4;
    float v[4];
    __device__ __forceinline__ void load(uint32_t addr) {
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
    }
};
template <> struct Vec16<double> {
    static constexpr int N = 2;
    double v[2];
    __device__ __forceinline__ void load(uint32_t addr) {
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(addr));
    }
};
// +-1 as T from the sign bit of a list word
__device__ __forceinline__ double word_sign_f64(uint32_t p) {
    return __hiloint2double((int) ((p & 0x80000000u) | 0x3ff00000u), 0);
}

// RL ("row per lane"): thread t of the CTA owns ROW t of the 1024 x 32 tile -- all 128 bytes of it, as eight 16-byte
// chunks -- instead of 16 bytes of eight rows. A warp then walks 32 lists in lockstep, one entry per lane and step:
// per step 8 address XORs, 8 LDS.128 and 32 FMAs serve up to 32 entries of 128 bytes (the group form spends 12 instructions
// per step on up to 4 entries, 16 bytes per lane), and the number of steps is the longest of the 32 lists (redux.sync.max).
// What bounds it is the shared-memory port: a quarter-warp costs a wavefront per LDS.128 as soon as one of its eight lanes has
// an entry at that step (lists are Poisson(2.75): 1.9 wavefronts issued per useful one). Lane l reads chunk j ^ (l & 7) in
// sub-step j, so the eight lanes of a quarter-warp hit eight different 16-byte bank groups whatever rows of Y they point at
// (conflict-free); accumulator set j of a lane therefore holds chunk j ^ (l & 7) of its row.
template <typename T, bool RL = false>
__global__ void __launch_bounds__(BN_THREADS, 1) saso_binned_kernel(const __grid_constant__ CUtensorMap tmY,
                                                                    const BinnedArgs<T> a) {
    constexpr int CW = BN_W * 4 / (int) sizeof(T);     // columns of C per CTA
    constexpr int VN = Vec16<T>::N;                    // columns per lane
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = tma::smem_u32(smem_raw);
    const uint32_t sbase = (raw + 127u) & ~127u;
    uint8_t* smem = smem_raw + (sbase - raw);
    const uint32_t bar_full = sbase + BN_OFF_BAR;          // 2 x 8 B
    const uint32_t bar_empty = bar_full + 16;              // 2 x 8 B
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + BN_OFF_BAR + 32);   // 2 counters

    const int tid = threadIdx.x, lane = tid & 31;
    const int gi = tid >> 3, l8 = tid & 7;
    const int col0 = (int) blockIdx.x * CW;
    const int64_t row0 = (int64_t) blockIdx.y * BN_PT;
    const int g = (int) blockIdx.z;
    const int nbox = a.Kc / BN_BOXR;
    const uint32_t ybytes = (uint32_t) a.Kc * BN_W * 4;

    if (tid == 0) {
        tma::mbar_init(bar_full, 1);
        tma::mbar_init(bar_full + 8, 1);
        tma::mbar_init(bar_empty, BN_THREADS / 32);
        tma::mbar_init(bar_empty + 8, BN_THREADS / 32);
        cnt[0] = 0;
        cnt[1] = 0;
        tma::mbar_fence_init();
    }
    __syncthreads();

    // one thread: Y slice + this row tile's part of the sorted list + its offsets -> stage b
    auto issue = [&](int c, int b) {
        const uint16_t* go = a.offs16 + (int64_t) c * a.Ppad + row0;
        const uint32_t lo = __ldg(go), hi = __ldg(go + BN_PT);
        const uint32_t lo4 = lo & ~3u, hi4 = (hi + 3u) & ~3u;
        const uint32_t lbytes = (hi4 - lo4) * 4u;
        const uint32_t bar = bar_full + 8u * (uint32_t) b;
        const uint32_t dst = sbase + (uint32_t) b * BN_STAGE;
        tma::mbar_arrive_expect_tx(bar, ybytes + lbytes + BN_OBYTES);
        for (int x = 0; x < nbox; ++x)
            tma::load_2d(dst + (uint32_t) x * (BN_BOXR * BN_W * 4), &tmY, bar, col0, c * a.Kc + x * BN_BOXR);   // 64 rows x 128 bytes
        if (lbytes) bulk_g2s(dst + BN_OFF_LIST, a.sorted + (int64_t) c * BN_ENT_CAP + lo4, lbytes, bar);
        bulk_g2s(dst + BN_OFF_OFFS, go, BN_OBYTES, bar);
    };

    T acc[BN_RPG][VN];
#pragma unroll
    for (int q = 0; q < BN_RPG; ++q)
#pragma unroll
        for (int e = 0; e < VN; ++e) acc[q][e] = (T) 0;

    if (tid == 0) {
        if (g < a.nchunks) issue(g, 0);
        if (g + a.G < a.nchunks) issue(g + a.G, 1);
    }

    int it = 0;
    for (int c = g; c < a.nchunks; c += a.G, ++it) {
        const int b = it & 1;
        const uint32_t par = (uint32_t) ((it >> 1) & 1);
        const uint32_t stage = sbase + (uint32_t) b * BN_STAGE;
        tma::mbar_wait(bar_full + 8u * (uint32_t) b, par);

        if constexpr (RL) {
            uint32_t o0, o1, lo;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(o0) : "r"(stage + BN_OFF_OFFS + 2u * (uint32_t) tid));
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(o1) : "r"(stage + BN_OFF_OFFS + 2u * (uint32_t) tid + 2u));
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(lo) : "r"(stage + BN_OFF_OFFS));
            const uint32_t len = o1 - o0;
            const uint32_t pa = stage + BN_OFF_LIST - 4u * (lo & ~3u) + 4u * o0;
            const uint32_t steps = __reduce_max_sync(0xffffffffu, len);
            const uint32_t yb = stage + (uint32_t) l8 * 16;
#pragma unroll 1
            for (uint32_t t = 0; t < steps; ++t) {
                if (t < len) {
                    const uint32_t w = lds32(pa + 4u * t);
                    const uint32_t e = yb + (w << 1);                 // row of Y (128-byte aligned) + this lane's rotation
                    if constexpr (sizeof(T) == 4) {
                        const float sg = word_sign(w);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            Vec16<T> y;
                            y.load(e ^ (uint32_t) (j << 4));
#pragma unroll
                            for (int x = 0; x < VN; ++x) acc[j][x] = fmaf(y.v[x], sg, acc[j][x]);
                        }
                    } else {
                        const double sg = word_sign_f64(w);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            Vec16<T> y;
                            y.load(e ^ (uint32_t) (j << 4));
#pragma unroll
                            for (int x = 0; x < VN; ++x) acc[j][x] = fma(y.v[x], sg, acc[j][x]);
                        }
                    }
                }
            }
        } else {
        // offsets of this group's 8 rows (relative to the chunk's list), 16-byte aligned in shared memory
        const uint32_t oaddr = stage + BN_OFF_OFFS + (uint32_t) gi * (BN_RPG * 2);
        uint32_t o01, o23, o45, o67;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(o01), "=r"(o23), "=r"(o45), "=r"(o67) : "r"(oaddr));
        uint32_t o8;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(o8) : "r"(oaddr + 16));
        uint32_t lo;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(lo) : "r"(stage + BN_OFF_OFFS));
        // list word e sits at lbase + 4 * e
        const uint32_t lbase = stage + BN_OFF_LIST - 4u * (lo & ~3u);
        uint32_t o[BN_RPG + 1];
        o[0] = o01 & 0xffffu; o[1] = o01 >> 16; o[2] = o23 & 0xffffu; o[3] = o23 >> 16;
        o[4] = o45 & 0xffffu; o[5] = o45 >> 16; o[6] = o67 & 0xffffu; o[7] = o67 >> 16; o[8] = o8;
        const uint32_t ybase = stage + (uint32_t) l8 * 16;
        // The lists of the group's 8 rows are back to back, so one pointer walks all of them and the NEXT list word is
        // always in flight while the current one is being used (one shared-memory latency per nonzero instead of two
        // on the dependent chain). The read one word past the end lands in the buffer's slack and is never used.
        uint32_t pa = lbase + 4u * o[0];
        uint32_t p = lds32(pa);
#pragma unroll
        for (int q = 0; q < BN_RPG; ++q) {
            const uint32_t pe = lbase + 4u * o[q + 1];
            // deliberately not unrolled: the lists are ~3 entries long, and the unrolled-by-2 form with its remainder
            // handling measured 5% slower (more code per row than work)
#pragma unroll 1
            for (; pa < pe; pa += 4) {
                const uint32_t pn = lds32(pa + 4);
                Vec16<T> y;
                y.load(ybase + (p << 1));
                if constexpr (sizeof(T) == 4) {
                    const float sg = word_sign(p);
#pragma unroll
                    for (int e = 0; e < VN; ++e) acc[q][e] = fmaf(y.v[e], sg, acc[q][e]);
                } else {
                    const double sg = word_sign_f64(p);
#pragma unroll
                    for (int e = 0; e < VN; ++e) acc[q][e] = fma(y.v[e], sg, acc[q][e]);
                }
                p = pn;
            }
        }
        }   // !RL
        // release the stage; the last warp to leave re-arms it with chunk c + 2G
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(bar_empty + 8u * (uint32_t) b);
            const uint32_t old = atomicAdd(&cnt[b], 1u);
            if ((old & (BN_THREADS / 32 - 1)) == (BN_THREADS / 32 - 1) && c + 2 * a.G < a.nchunks) {
                tma::mbar_wait(bar_empty + 8u * (uint32_t) b, par);
                issue(c + 2 * a.G, b);
            }
        }
    }

    // ---- add the tile into C (beta was applied beforehand; G CTAs share the tile) ----
#pragma unroll
    for (int q = 0; q < BN_RPG; ++q) {
        // group form: lane = 16 bytes of rows 8 gi .. 8 gi + 7; row-per-lane form: accumulator set q = chunk q ^ l8 of row tid
        const int64_t col = (int64_t) col0 + (RL ? (q ^ l8) : l8) * VN;
        const int64_t row = RL ? row0 + tid : row0 + (int64_t) gi * BN_RPG + q;
        if (row >= a.P || col >= a.Q) continue;
        T* cp = a.C + row * a.crs + col;
        if constexpr (sizeof(T) == 4) {
            const float v0 = a.alpha * acc[q][0], v1 = a.alpha * acc[q][1], v2 = a.alpha * acc[q][2], v3 = a.alpha * acc[q][3];
            if (a.c_vec4 && col + 3 < a.Q) {
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(cp), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
            } else {
                atomicAdd(cp, v0);
                if (col + 1 < a.Q) atomicAdd(cp + 1, v1);
                if (col + 2 < a.Q) atomicAdd(cp + 2, v2);
                if (col + 3 < a.Q) atomicAdd(cp + 3, v3);
            }
        } else {
            // (red has no vector form for f64: ptxas "Vector qualifier is not allowed")
            atomicAdd(cp, a.alpha * acc[q][0]);
            if (col + 1 < a.Q) atomicAdd(cp + 1, a.alpha * acc[q][1]);
        }
    }
}

template <int G, bool VEC = false, int NT = BIN_THREADS, int MINB = 1>
int launch_bin(const BinArgs& a, size_t smem, cudaStream_t st) {
    static DevOnce attr_done;
    if (attr_done.need()) {
        if (cudaFuncSetAttribute(saso_bin_kernel<G, VEC, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
        attr_done.done();
    }
    int64_t grid = a.nchunks;
    const int64_t cap = (int64_t) sm_count() * 4;
    if (grid > cap) grid = cap;
    saso_bin_kernel<G, VEC, NT, MINB><<<(unsigned) grid, NT, smem, st>>>(a);
    return 0;
}

}  // namespace

// Returns 0 if the product was computed, -1 if this path does not take the problem (caller falls back), >0 on error.
// C must have been beta-scaled by the caller.
template <typename T>
int launch_saso_binned(const SasoProblem<T>& p, cudaStream_t st) {
    constexpr int CW = BN_W * 4 / (int) sizeof(T);     // columns of C per CTA
    const int64_t path = get_option("saso_path");
    if (path == 1) return -1;                                 // 1 = force the atomic kernel
    const bool scatter = p.major_is_rows ? !p.x_is_transposed : p.x_is_transposed;   // short axis <-> rows of C
    if (!scatter) return -1;
    if (p.ycs != 1 || p.ccs != 1) return -1;
    if ((reinterpret_cast<uintptr_t>(p.Y) & 15) != 0 || ((p.yrs * (int64_t) sizeof(T)) & 15) != 0) return -1;   // TMA alignment rules
    if (p.vec_nnz > 32 || p.dim_major >= 0x7fffffffLL || p.P > BN_PMAX) return -1;
    if (p.Q > 0x7fffffffLL || p.yrs > 0x3fffffffLL) return -1;
    const int64_t w0 = p.major_is_rows ? p.co_s : p.ro_s;      // first column of X in operator coordinates
    const int64_t m0 = p.major_is_rows ? p.ro_s : p.co_s;      // first row of X
    const int64_t nvec = p.K;
    if (path != 2 && nvec * p.vec_nnz < 32768) return -1;      // small: one launch of the atomic kernel wins
    tma::EncodeTiledFn enc = tma::encode_tiled_fn();
    if (!enc) return -1;
    {
        static DevOnce attr_done;
        if (attr_done.need()) {
            if (cudaFuncSetAttribute(saso_binned_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BN_SMEM) != cudaSuccess ||
                cudaFuncSetAttribute(saso_binned_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BN_SMEM) != cudaSuccess) {
                cudaGetLastError();
                return -1;
            }
            attr_done.done();
        }
    }
    const int k = (int) p.vec_nnz;
    int Kc = (BN_ENT / k) / BN_BOXR * BN_BOXR;
    if (Kc > BN_KMAX) Kc = BN_KMAX;
    const int64_t ns = (p.Q + CW - 1) / CW, np = (p.P + BN_PT - 1) / BN_PT;
    if (ns > 0x7fffffffLL || np > 65535) return -1;
    const int Prows = (int) np * BN_PT, Ppad = Prows + 8;
    const size_t bin_smem = ((size_t) Prows + 32 + 2 * BN_ENT) * 4;
    const int sms = sm_count();
    // segments bound the workspace (sorted list + offsets) to about 1 GiB
    const int64_t per_chunk = (int64_t) BN_ENT_CAP * 4 + (int64_t) Ppad * 2;
    int64_t seg_chunks = ((int64_t) 1 << 30) / per_chunk;
    if (seg_chunks < 1) seg_chunks = 1;
    const int64_t seg_vecs = seg_chunks * Kc;
    for (int64_t v0 = 0; v0 < nvec; v0 += seg_vecs) {
        const int64_t nv = (nvec - v0 < seg_vecs) ? nvec - v0 : seg_vecs;
        const int64_t nchunks = (nv + Kc - 1) / Kc;
        uint32_t* sorted = (uint32_t*) workspace(7, (size_t) nchunks * BN_ENT_CAP * 4, st);
        uint16_t* offs16 = (uint16_t*) workspace(6, (size_t) nchunks * Ppad * 2 + 64, st);
        if (!sorted || !offs16) return fail_cuda(cudaErrorMemoryAllocation, "SASO binning workspace");

        BinArgs b;
        b.ctr = p.ctr; b.key = p.key; b.k = k; b.dim_major = (uint32_t) p.dim_major;
        b.vec_lo = w0 + v0; b.nvec = nv; b.m0 = m0; b.P = p.P; b.Kc = Kc; b.nchunks = nchunks;
        b.Prows = Prows; b.Ppad = Ppad; b.sorted = sorted; b.offs16 = offs16;
        for (int j = 0; j < 16; ++j) b.magic[j] = (j < k && p.dim_major - j > 0) ? 0xffffffffu / (uint32_t) (p.dim_major - j) : 0u;
        // saso_bin_path: 0 auto (thread per column for vec_nnz = 2, 4, 8, 16), 1 lane per entry always
        const bool vec = get_option("saso_bin_path") == 0 && p.dim_major >= k;
        int rc;
        // (352 threads x 3 CTAs per SM -- two FULL passes over a 704-column chunk -- measured slower than 512 x 2 with a
        // ragged second pass: 71.0 against 64.5 us per 1e6 columns; the template parameters stay for such experiments)
        if (vec && k == 2) rc = launch_bin<2, true, BIN_THREADS, 2>(b, bin_smem, st);
        else if (vec && k == 4) rc = launch_bin<4, true, BIN_THREADS, 2>(b, bin_smem, st);
        else if (vec && k == 8) rc = launch_bin<8, true, BIN_THREADS, 2>(b, bin_smem, st);
        else if (vec && k == 16) rc = launch_bin<16, true, BIN_THREADS, 1>(b, bin_smem, st);
        else if (k <= 1) rc = launch_bin<1>(b, bin_smem, st);
        else if (k <= 2) rc = launch_bin<2>(b, bin_smem, st);
        else if (k <= 4) rc = launch_bin<4>(b, bin_smem, st);
        else if (k <= 8) rc = launch_bin<8>(b, bin_smem, st);
        else if (k <= 16) rc = launch_bin<16>(b, bin_smem, st);
        else rc = launch_bin<32>(b, bin_smem, st);
        if (rc) return v0 == 0 ? -1 : fail("SASO binning kernel could not be configured");
        count_launch();
        RB_CUDA(cudaGetLastError());

        CUtensorMap tm;
        const cuuint64_t gdim[2] = {(cuuint64_t) p.Q, (cuuint64_t) nv};
        const cuuint64_t gstr[1] = {(cuuint64_t) p.yrs * sizeof(T)};
        const cuuint32_t box[2] = {(cuuint32_t) CW, BN_BOXR};
        const cuuint32_t estr[2] = {1, 1};
        CUresult cr = enc(&tm, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2,
                          const_cast<T*>(p.Y + v0 * p.yrs), gdim, gstr, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return v0 == 0 ? -1 : fail("cuTensorMapEncodeTiled failed for the SASO apply");

        BinnedArgs<T> a;
        a.sorted = sorted; a.offs16 = offs16; a.nchunks = (int) nchunks; a.Kc = Kc; a.Ppad = Ppad;
        int64_t G = sms / (ns * np);
        if (G < 1) G = 1;
        if (G > nchunks) G = nchunks;
        if (G > 65535) G = 65535;
        a.G = (int) G;
        a.P = p.P; a.Q = p.Q;
        a.alpha = p.alpha;
        a.C = p.C; a.crs = p.crs;
        a.c_vec4 = ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && ((p.crs * (int64_t) sizeof(T)) & 15) == 0) ? 1 : 0;
        dim3 grid((unsigned) ns, (unsigned) np, (unsigned) G);
        // saso_rows: 1 = a lane owns a whole row of the tile (RL), 0 = an 8-lane group owns 8 rows
        if (get_option("saso_rows") != 0) saso_binned_kernel<T, true><<<grid, BN_THREADS, BN_SMEM, st>>>(tm, a);
        else saso_binned_kernel<T, false><<<grid, BN_THREADS, BN_SMEM, st>>>(tm, a);
        count_launch();
        count_owner_launch();
        RB_CUDA(cudaGetLastError());
    }
    return 0;
}

template int launch_saso_binned<float>(const SasoProblem<float>&, cudaStream_t);
template int launch_saso_binned<double>(const SasoProblem<double>&, cudaStream_t);

}  // namespace rb
