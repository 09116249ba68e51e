// K1: fill_dense -- a window of a DenseSkOp sample written straight into the requested layout.
//
// Replaces (reference file:line): dense::fill_dense_submat_impl (RandBLAS/dense_skops.hh:96-170),
// the Uniform post-scale (:587-590) and the out-of-place layout flip (:596-604) of fill_dense_unpacked.
//
// Work unit = one Philox4x32-10 counter block (4 samples) per thread per step. In the operator's
// natural layout the sample is dim_minor "major-axis vectors" of R = ceil(dim_major/4) blocks each;
// element (v, u) is lane u%4 of block (seed + v*R + u/4). The window is [v0, v0+nv) x [u0, u0+nu).
// Threads walk the linear block index L = v_local * nblk + b with a grid stride, so consecutive
// lanes write consecutive 16 B (float) / 32 B (double) pieces of the same vector: fully coalesced
// 128-bit / 256-bit stores when the destination is written in natural orientation. When the
// requested layout is the transpose of the natural one the lanes of a warp walk v instead, so each
// of the four per-thread stores is still a contiguous warp-wide segment.
//
// Roofline: HBM write. Algorithmic bytes = sizeof(T) per sample.
#include "common.cuh"
#include "kernels.h"

namespace rb {

struct FillArgs {
    Ctr128 ctr;
    PhiloxKey key;
    int64_t R;          // blocks per major-axis vector of the parent operator
    int64_t v0, nv;     // vector window
    int64_t u0, nu;     // position window inside a vector
    int64_t blk_first;  // u0 / 4
    int64_t nblk;       // blocks touched per vector
    int64_t sv, su;     // destination strides (elements) per vector step / per position step
    int64_t total;      // nv * nblk
    int64_t q_step, r_step;  // grid stride decomposed: stride = q_step * nblk + r_step
    const double2* logtab;
};

template <typename T>
__device__ __forceinline__ void store4_vec(T* p, T a, T b, T c, T d);
template <>
__device__ __forceinline__ void store4_vec<float>(float* p, float a, float b, float c, float d) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <>
__device__ __forceinline__ void store4_vec<double>(double* p, double a, double b, double c, double d) {
    // 256-bit store (sm_100+): one instruction per Philox block of doubles
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <typename T, bool GAUSS, bool WALK_V>
__global__ void __launch_bounds__(256) fill_dense_kernel(const FillArgs a, T* __restrict__ dst) {
    __shared__ __align__(16) double2 logtab[GAUSS ? LOGF_TABLE_ENTRIES : 1];
    if constexpr (GAUSS) {
        load_logf_table(logtab, a.logtab);
        __syncthreads();
    }
    int64_t L = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= a.total) return;
    // (vl, b): local vector index and block index inside the window
    int64_t vl, b;
    if constexpr (WALK_V) {          // lanes walk vectors: L = b * nv + vl
        b = L / a.nv;
        vl = L - b * a.nv;
    } else {                         // lanes walk blocks of one vector: L = vl * nblk + b
        vl = L / a.nblk;
        b = L - vl * a.nblk;
    }
    const int64_t inner = WALK_V ? a.nv : a.nblk;
    for (; L < a.total; L += (int64_t) gridDim.x * blockDim.x) {
        const int64_t blk = a.blk_first + b;
        const Ctr128 c = ctr_add(a.ctr, (uint64_t)((a.v0 + vl) * a.R + blk));
        const float4 f = transform4<GAUSS>(philox4x32_10(c, a.key), logtab);
        const T x0 = finish_sample<T, GAUSS>(f.x), x1 = finish_sample<T, GAUSS>(f.y),
                x2 = finish_sample<T, GAUSS>(f.z), x3 = finish_sample<T, GAUSS>(f.w);
        const int64_t ur = blk * 4 - a.u0;                 // position of lane 0 relative to the window
        T* p = dst + vl * a.sv + ur * a.su;
        const bool full = (ur >= 0) && (ur + 4 <= a.nu);
        if (!WALK_V && full && a.su == 1 && ((reinterpret_cast<uintptr_t>(p) & (4 * sizeof(T) - 1)) == 0)) {
            store4_vec<T>(p, x0, x1, x2, x3);
        } else {
            if (ur + 0 >= 0 && ur + 0 < a.nu) p[0] = x0;
            if (ur + 1 >= 0 && ur + 1 < a.nu) p[1 * a.su] = x1;
            if (ur + 2 >= 0 && ur + 2 < a.nu) p[2 * a.su] = x2;
            if (ur + 3 >= 0 && ur + 3 < a.nu) p[3 * a.su] = x3;
        }
        // advance (vl, b) by the grid stride without a division
        if constexpr (WALK_V) {
            vl += a.r_step; b += a.q_step;
            if (vl >= inner) { vl -= inner; b += 1; }
        } else {
            b += a.r_step; vl += a.q_step;
            if (b >= inner) { b -= inner; vl += 1; }
        }
    }
}

// Fast path for long vectors: a tile is FILL_UNROLL x 256 consecutive blocks of ONE vector, so the counter of
// a block is (tile-uniform 128-bit base) + (small offset) and the destination pointer is base + 4 * offset.
// Tiles whose blocks are all interior and 4-element aligned take a predicate-free store path.
constexpr int FILL_UNROLL = 4;   // Uniform family; the kernel is templated on the unroll and on the CTAs per SM

struct FillTileArgs {
    Ctr128 ctr;
    PhiloxKey key;
    int64_t R, v0, nv, u0, nu, blk_first, nblk, sv;
    int64_t tiles_per_vec, total_tiles;
    int64_t q_step, r_step;     // gridDim.x = q_step * tiles_per_vec + r_step
    const double2* logtab;
};

// REP: a tile is REP passes of FILL_UNROLL x 256 blocks. The per-tile set-up (128-bit counter base, 64-bit tile / pointer
// arithmetic, interior test: ~87 instructions, every thread) is then shared by REP x FILL_UNROLL blocks without the register
// pressure of a longer unrolled body.
template <typename T, bool GAUSS, int FILL_UNROLL = 4, int MINB = 4, int REP = 1>
__global__ void __launch_bounds__(256, MINB) fill_dense_tiled_kernel(const FillTileArgs a, T* __restrict__ dst) {
    __shared__ __align__(16) double2 logtab[GAUSS ? LOGF_TABLE_ENTRIES : 1];
    if constexpr (GAUSS) {
        load_logf_table(logtab, a.logtab);
        __syncthreads();
    }
    int64_t t = blockIdx.x;
    if (t >= a.total_tiles) return;
    int64_t vl = t / a.tiles_per_vec;
    int64_t ch = t - vl * a.tiles_per_vec;
    constexpr int64_t TILE = 256 * FILL_UNROLL * REP;
    for (; t < a.total_tiles; t += gridDim.x) {
        const int64_t b0 = ch * TILE;                                  // first block of the tile (window-relative)
        const Ctr128 base = ctr_add(a.ctr, (uint64_t) ((a.v0 + vl) * a.R + a.blk_first + b0));
        const uint64_t base_lo = ((uint64_t) base.c1 << 32) | base.c0;
        const int64_t ur0 = (a.blk_first + b0) * 4 - a.u0;             // position of the tile's first sample
        T* p0 = dst + vl * a.sv + ur0;
        const int64_t nb = min((int64_t) TILE, a.nblk - b0);           // blocks in this tile
        const bool interior = (ur0 >= 0) && (ur0 + 4 * nb <= a.nu) &&
                              ((reinterpret_cast<uintptr_t>(p0) & (4 * sizeof(T) - 1)) == 0);
        // predicate-free path: whole interior tile whose block counters do not carry out of the low 64 bits
        if (interior && nb == TILE && base_lo + (uint64_t) TILE >= base_lo) {
#pragma unroll 1
            for (int rep = 0; rep < REP; ++rep) {
#pragma unroll
                for (int j = 0; j < FILL_UNROLL; ++j) {
                    const uint32_t off = (uint32_t) (rep * FILL_UNROLL + j) * 256u + threadIdx.x;
                    const uint64_t lo = base_lo + off;
                    const Ctr128 c{(uint32_t) lo, (uint32_t) (lo >> 32), base.c2, base.c3};
                    const float4 f = transform4<GAUSS>(philox4x32_10(c, a.key), logtab);
                    store4_vec<T>(p0 + 4 * off, finish_sample<T, GAUSS>(f.x), finish_sample<T, GAUSS>(f.y),
                                  finish_sample<T, GAUSS>(f.z), finish_sample<T, GAUSS>(f.w));
                }
            }
        } else {
            for (int j = 0; j < FILL_UNROLL * REP; ++j) {
                const int64_t off = j * 256 + threadIdx.x;
                if (off >= nb) break;
                const Ctr128 c = ctr_add(base, (uint64_t) off);
                const float4 f = transform4<GAUSS>(philox4x32_10(c, a.key), logtab);
                const T x0 = finish_sample<T, GAUSS>(f.x), x1 = finish_sample<T, GAUSS>(f.y),
                        x2 = finish_sample<T, GAUSS>(f.z), x3 = finish_sample<T, GAUSS>(f.w);
                const int64_t ur = ur0 + 4 * off;
                T* p = p0 + 4 * off;
                if (ur >= 0 && ur + 4 <= a.nu && ((reinterpret_cast<uintptr_t>(p) & (4 * sizeof(T) - 1)) == 0)) {
                    store4_vec<T>(p, x0, x1, x2, x3);
                } else {
                    if (ur + 0 >= 0 && ur + 0 < a.nu) p[0] = x0;
                    if (ur + 1 >= 0 && ur + 1 < a.nu) p[1] = x1;
                    if (ur + 2 >= 0 && ur + 2 < a.nu) p[2] = x2;
                    if (ur + 3 >= 0 && ur + 3 < a.nu) p[3] = x3;
                }
            }
        }
        ch += a.r_step; vl += a.q_step;
        if (ch >= a.tiles_per_vec) { ch -= a.tiles_per_vec; vl += 1; }
    }
}

__global__ void philox_words_kernel(Ctr128 ctr, PhiloxKey key, int64_t n_blocks, uint4* __restrict__ out) {
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_blocks;
         i += (int64_t) gridDim.x * blockDim.x)
        out[i] = philox4x32_10(ctr_add(ctr, (uint64_t) i), key);
}

__global__ void __launch_bounds__(256) boxmuller_words_kernel(int64_t n, const uint32_t* __restrict__ w0,
                                                              const uint32_t* __restrict__ w1, float* __restrict__ g0,
                                                              float* __restrict__ g1, const double2* __restrict__ gtab) {
    __shared__ __align__(16) double2 logtab[LOGF_TABLE_ENTRIES];
    load_logf_table(logtab, gtab);
    __syncthreads();
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
        float a, b;
        boxmuller(w0[i], w1[i], logtab, a, b);
        g0[i] = a;
        g1[i] = b;
    }
}

int launch_boxmuller_words(int64_t n, const uint32_t* w0, const uint32_t* w1, float* g0, float* g1, cudaStream_t st) {
    if (n <= 0) return 0;
    const double2* tab = logf_table_device();
    if (!tab) return fail_cuda(cudaErrorMemoryAllocation, "logf table");
    int64_t grid = (n + 255) / 256;
    if (grid > (int64_t) sm_count() * 8) grid = (int64_t) sm_count() * 8;
    boxmuller_words_kernel<<<(unsigned) grid, 256, 0, st>>>(n, w0, w1, g0, g1, tab);
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}

int launch_philox_words(Ctr128 ctr, PhiloxKey key, int64_t n_blocks, uint32_t* out, cudaStream_t st) {
    if (n_blocks <= 0) return 0;
    int64_t grid = (n_blocks + 255) / 256;
    int64_t cap = (int64_t) sm_count() * 16;
    if (grid > cap) grid = cap;
    philox_words_kernel<<<(unsigned) grid, 256, 0, st>>>(ctr, key, n_blocks, reinterpret_cast<uint4*>(out));
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}

// Device-pointer launcher. (v0,nv,u0,nu) is the window in natural coordinates; (sv,su) the
// destination strides for one vector step / one position step.
template <typename T>
int launch_fill_dense(const DenseGen& g, char family, int64_t v0, int64_t nv, int64_t u0, int64_t nu, T* dst,
                      int64_t sv, int64_t su, cudaStream_t st) {
    if (nv <= 0 || nu <= 0) return 0;
    if (family == 'G' && !g.logtab) return fail_cuda(cudaErrorMemoryAllocation, "logf table");
    FillArgs a;
    a.ctr = g.ctr; a.key = g.key; a.R = g.R; a.logtab = g.logtab;
    a.v0 = v0; a.nv = nv; a.u0 = u0; a.nu = nu;
    a.blk_first = u0 / 4;
    a.nblk = (u0 + nu - 1) / 4 - a.blk_first + 1;
    a.sv = sv; a.su = su;
    a.total = nv * a.nblk;
    // lanes walk whichever direction is contiguous in memory
    const bool walk_v = (su != 1) && (sv == 1 || sv < su);
    const bool gauss_ = family == 'G';
    // Tiled fast path. Gaussian: 8 blocks per thread and tile, 5 CTAs (10 warps per scheduler, 48 registers) -- the
    // Box-Muller chains want both the instruction-level and the thread-level parallelism (measured on C2: 561 vs 518
    // Gsamples/s for 4 blocks / 4 CTAs; tools/exp_fill_variants.py). Uniform: 4 blocks, 4 CTAs (already at the HBM roofline).
    // Uniform float, long vectors: 16 blocks per thread and tile, 3 CTAs per SM. With 16 bytes written per Philox block this
    // instantiation is bound by instructions, not by HBM, and the per-tile set-up (128-bit counter base, tile index
    // arithmetic, interior test: ~87 instructions) is a quarter of the work at 4 blocks of 61 instructions. Measured
    // (2048 x 1e6 window, tools/exp_fill_unroll.py): 4 blocks / 4 CTAs 1157 Gsamples/s, 8 / 4 1289, 8 / 5 1282, 16 / 4 1346,
    // 16 / 3 1367 = 5.47 TB/s = 0.84 of measured HBM; bit-identical. "fill_unroll" = 0 restores 4 blocks.
    const bool u16 = !gauss_ && sizeof(T) == 4 && a.nblk >= 2 * 256 * 16 && get_option("fill_unroll") != 0;
    // Gaussian, long vectors: tiles of REP = 4 passes of 8 blocks ("fill_rep" = 0: one pass, 2: two passes). Measured
    // (tools/exp_fill_rep.py, 1024 x 1e6 window): 558 -> 570 Gsamples/s (two passes: 551, the tile count no longer divides
    // the grid well); the full C2 fill 553 -> 563; bit-identical.
    const int64_t ropt = get_option("fill_rep");
    const bool rep4 = gauss_ && a.nblk >= 2 * 256 * 8 * 4 && ropt == 1;
    const bool rep2 = gauss_ && a.nblk >= 2 * 256 * 8 * 2 && ropt == 2;
    const int unroll = gauss_ ? (rep4 ? 32 : rep2 ? 16 : 8) : (u16 ? 16 : FILL_UNROLL);       // blocks per thread and tile
    if (!walk_v && su == 1 && a.nblk >= 2 * 256 * unroll) {
        FillTileArgs t;
        t.ctr = g.ctr; t.key = g.key; t.R = g.R; t.logtab = g.logtab; t.v0 = v0; t.nv = nv; t.u0 = u0; t.nu = nu;
        t.blk_first = a.blk_first; t.nblk = a.nblk; t.sv = sv;
        t.tiles_per_vec = (a.nblk + 256 * unroll - 1) / (256 * unroll);
        t.total_tiles = nv * t.tiles_per_vec;
        int64_t tgrid = t.total_tiles;
        const int64_t tcap = (int64_t) sm_count() * (gauss_ ? 10 : 8);
        if (tgrid > tcap) tgrid = tcap;
        t.q_step = tgrid / t.tiles_per_vec;
        t.r_step = tgrid % t.tiles_per_vec;
        if (rep4) fill_dense_tiled_kernel<T, true, 8, 5, 4><<<(unsigned) tgrid, 256, 0, st>>>(t, dst);
        else if (rep2) fill_dense_tiled_kernel<T, true, 8, 5, 2><<<(unsigned) tgrid, 256, 0, st>>>(t, dst);
        else if (gauss_) fill_dense_tiled_kernel<T, true, 8, 5><<<(unsigned) tgrid, 256, 0, st>>>(t, dst);
        else if (u16) fill_dense_tiled_kernel<T, false, 16, 3><<<(unsigned) tgrid, 256, 0, st>>>(t, dst);
        else fill_dense_tiled_kernel<T, false, FILL_UNROLL, 4><<<(unsigned) tgrid, 256, 0, st>>>(t, dst);
        count_launch();
        RB_CUDA(cudaGetLastError());
        return 0;
    }
    int64_t grid = (a.total + 255) / 256;
    const int64_t cap = (int64_t) sm_count() * 8;      // 8 CTAs of 256 threads per SM, then grid-stride
    if (grid > cap) grid = cap;
    const int64_t stride = grid * 256;
    const int64_t inner = walk_v ? nv : a.nblk;
    a.q_step = stride / inner;
    a.r_step = stride % inner;
    const bool gauss = family == 'G';
#define RB_LAUNCH_FILL(G, W) fill_dense_kernel<T, G, W><<<(unsigned) grid, 256, 0, st>>>(a, dst)
    if (gauss) { if (walk_v) RB_LAUNCH_FILL(true, true); else RB_LAUNCH_FILL(true, false); }
    else       { if (walk_v) RB_LAUNCH_FILL(false, true); else RB_LAUNCH_FILL(false, false); }
#undef RB_LAUNCH_FILL
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}

template int launch_fill_dense<float>(const DenseGen&, char, int64_t, int64_t, int64_t, int64_t, float*, int64_t,
                                      int64_t, cudaStream_t);
template int launch_fill_dense<double>(const DenseGen&, char, int64_t, int64_t, int64_t, int64_t, double*, int64_t,
                                       int64_t, cudaStream_t);

}  // namespace rb
