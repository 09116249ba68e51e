// K2a: fused generate-and-multiply dense sketch for float on the 5th-generation tensor cores (3xTF32).
//
// Replaces dense::lskge3 / rskge3 (RandBLAS/skge.hh:154-202, 307-355) for float when the operator is not
// materialised: the reference first fills a d x m host buffer (submatrix_as_blackbox, dense_skops.hh:677-688)
// and then calls blas::gemm (skge.hh:200); here each 128 x 32 tile of S is regenerated from
// (key, counter, ro_s, co_s) straight into shared memory in the K-major 128B-swizzled layout tcgen05 reads,
// already split into its two TF32 terms, so S never touches HBM.
//
// Canonical problem (see kernels.h): C(P x Q) = alpha * X(P x K) * Y(K x Q) + beta * C with X = op(S window).
// One CTA owns a 128 x 256 tile of C and one K range (split-K over blockIdx.z):
//   warp 0      : TMA producer. Y tiles (256 columns x 32 k, K contiguous) land in shared memory, 128B swizzle.
//   warp 1      : allocates all 512 TMEM columns; one thread issues tcgen05.mma kind::tf32, cta_group::1, M=128,
//                 N=256, K=8: per 8-deep slice three MMAs  X_lo*Y_hi + X_hi*Y_lo -> accumulator "small",
//                 X_hi*Y_hi -> accumulator "big"  (3xTF32: fp32 operands are split x = hi + lo with hi, lo exactly
//                 representable in TF32; the dropped lo*lo term is 2^-22 relative). Two 128 x 256 fp32
//                 accumulators live in TMEM: the tensor core adds into its accumulator with truncation, a bias that
//                 grows with the number of MMAs added, so the cross terms (whose rounding does not matter) are
//                 kept out of the accumulator that carries the leading term.
//   warps 2..17 : generators. Per K step: Philox4x32-10 + uneg11/Box-Muller for the X tile -> (hi, lo) ->
//                 swizzled st.shared; then Y_lo = Y - trunc_tf32(Y) for the tile TMA just delivered (the tensor
//                 core ignores the 13 low mantissa bits of its 32-bit operands, so the raw tile IS Y_hi).
//                 After the last K step the same warps run the epilogue: tcgen05.ld -> alpha/beta -> global.
// Pipeline: 2 stages of {X_hi, X_lo, Y, Y_lo} = 96 KB each; mbarriers full_y (TMA -> consumers),
// ready (generators -> MMA), empty (tcgen05.commit -> producers), accum (last commit -> epilogue).
// No partial sum stays in TMEM for more than MAX_CHAIN_STEPS K steps (accuracy, see launch_dense_tc_f32);
// split-K partial tiles go to a workspace and are summed in a fixed order, in round-to-nearest fp32, by a
// second small kernel, so the result does not depend on the grid (the reference guarantees thread-count invariance, dense_skops.hh:90-94).
//
// Data contiguous along Q instead of K (left sketch of RowMajor data, right sketch of ColMajor data -- the
// range-finder call A * S): the Y tile arrives un-swizzled as 32 k-rows of 256 q (one TMA box) in a raw staging
// buffer and the generator warps transpose it into the K-major swizzled Y / Y_lo tiles (4 conflict-free 4-byte reads,
// 2 conflict-free 16-byte writes per 4 elements) before they generate X, so the next raw tile travels while X is
// being generated. The MMA side is unchanged. (Feeding the Q-contiguous tile to tcgen05 directly as an MN-major
// operand -- instruction-descriptor bit 16, LBO 4096 / SBO 1024 -- faulted with "illegal memory access" on this
// toolchain and was dropped.)
//
// Materialised operators contiguous along the ROWS of X (a filled Axis::Short operator): the tile is an MN-major A operand
// (instruction-descriptor bit 15), delivered by four TMA boxes of 32 rows x 32 k with the same 32-byte-atom swizzle.
//
// Roofline: tensor. 2*P*Q*K algorithmic flops, 3 MMAs issued per product => peak = TF32 dense peak / 3.
#include <cuda.h>
#include "common.cuh"
#include "kernels.h"

namespace rb {

namespace {

constexpr int BM = 128, BN = 256, BK = 32, STAGES = 2;
constexpr int GEN_WARPS = 16;            // 4 per scheduler: the Philox chains need the thread-level parallelism
constexpr int X_PER_THREAD = (BM * BK / 4) / (32 * GEN_WARPS);   // Philox blocks per generator thread per K step
constexpr int TC_THREADS = 64 + 32 * GEN_WARPS;
constexpr uint32_t X_BYTES = BM * BK * 4, Y_BYTES = BN * BK * 4;
constexpr uint32_t STAGE_BYTES = 2 * X_BYTES + 2 * Y_BYTES;
constexpr int RAW_K = 16;                                // k-rows per raw tile: the 32-deep step arrives in two halves
constexpr uint32_t RAW_BYTES = RAW_K * BN * 4;
constexpr uint32_t RAW_OFFSET = STAGES * STAGE_BYTES;   // raw Q-contiguous Y tile (y_mn mode): RAW_K k-rows x 256 q
constexpr uint32_t BAR_OFFSET = RAW_OFFSET + RAW_BYTES;
constexpr uint32_t TC_SMEM = BAR_OFFSET + 128 + 1024;   // barriers + TMEM slot + 1 KB alignment slack
static_assert(TC_SMEM + 8752 + 64 <= 232448, "shared memory budget (dynamic + the Gaussian table)");
static_assert(STAGES == 2, "the 2-CTA cluster mode maps stage s to generating CTA s");
constexpr uint32_t TMEM_COLS = 512;      // two 128 x 256 fp32 accumulators: hi*hi and the two cross terms
constexpr int MAX_CHAIN_STEPS = 40;       // K steps accumulated in TMEM before the partial sum leaves the tensor core

struct TcArgs {
    Ctr128 ctr;
    PhiloxKey key;
    const double2* logtab;
    int64_t R;
    int64_t v0;        // first operator vector (row of X)
    int64_t ublk0;     // Philox block that holds position k = 0 of a vector: u0 / 4
    int kshift;        // u0 & 3: when non-zero every 4-wide chunk of X straddles two Philox blocks
    int64_t P, Q;
    int steps_total;   // ceil(K / 32)
    int y_mn;          // Y is Q-contiguous: 2 = its tiles go to the tensor core as they are (MN-major B operand, TMA boxes of
                       // 32 q x 32 k with the 32-byte-atom 128B swizzle); 1 = raw tiles transposed by the generator warps (round 2's
                       // first form, "tc_ymn" = 1); 0 = Y is K-contiguous
    int x_t;           // 1: the operator's Philox blocks run along the ROWS of X (Axis::Short operators, transposed use of a
                       //    Long one): a block is 4 consecutive rows of one column k; v0 then counts along k, ublk0 along i
    int splits;
    float alpha, beta;
    float* C;
    int64_t crs, ccs;
    int64_t xr0;       // XMAT: first row of X in the materialised operator (tensor-map coordinates)
    int64_t xk0;       // XMAT: first column of X
    int x_mn;          // XMAT: the materialised operator is contiguous along the ROWS of X (a filled Axis::Short operator): its
                       // tiles go to the tensor core as an MN-major A operand, 4 TMA boxes of 32 rows x 32 k (32-byte-atom swizzle)
    float* W;          // split-K workspace: W[split][j][i], i fastest, ld = P_pad; null when splits == 1
    int64_t P_pad, Q_pad;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
// Bounded wait: a pipeline bug must end in a trap (a launch error the host sees), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if ((spins & 0xfff) == 0xfff) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) __trap();   // ~2 s at 1.9 GHz
        }
    }
}
// Same bounded wait with acquire semantics at cluster scope: the data behind the barrier was written by the peer CTA.
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if ((spins & 0xfff) == 0xfff) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) __trap();
        }
    }
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, const float4& v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// all state spaces: the generic-proxy writes to order include stores into the peer CTA's shared memory
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

// L2 prefetch of a tile that will be loaded a few steps later: with only two shared-memory stages the HBM
// latency of the actual load would otherwise sit on the critical path of every K step.
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}

// K-major, 128B-swizzled shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
    return (uint64_t) ((addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t) (1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major (contiguous along N) B operand, tf32: the only shared-memory layout the tensor core takes is the 128B swizzle with
// 32-byte atoms (cute: SWIZZLE_128B_BASE32B, Swizzle<2,5,2>; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 128 bytes = 32
// consecutive columns of one k, the 32-byte chunk index XORed with (k & 3), atoms of 4 k rows. Leading byte offset = distance
// between two 32-column blocks (one TMA box of 32 k rows: 4096), stride byte offset = distance between the two 4-row atoms of a
// K = 8 instruction (512). Probed in isolation first: tools/micro/mn_major_probe.cu (the plain 128B swizzle returns garbage).
__device__ __forceinline__ uint64_t smem_desc_mn32(uint32_t addr) {
    return (uint64_t) ((addr >> 4) & 0x3fffu) | ((uint64_t) (4096u >> 4) << 16) | ((uint64_t) (512u >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
// kind::tf32, fp32 accumulate, both operands K-major, M = 128, N = 256; IDESC_B_MN: the B operand is MN-major
constexpr uint32_t IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (BN >> 3) << 17) | ((uint32_t) (BM >> 4) << 24);
constexpr uint32_t IDESC_B_MN = 1u << 16;
constexpr uint32_t IDESC_A_MN = 1u << 15;   // the A operand is MN-major (same 32-byte-atom layout, blocks of 32 rows)

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t accumulate, uint32_t idesc = IDESC_TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- CTA pair (cta_group::2): the two SMs of a TPC execute ONE M = 256 MMA; each holds its 128 rows of X and of the
// accumulator and HALF of the Y tile (128 of the 256 columns), so per SM and K step the tensor core reads half the Y
// bytes, TMA writes half and the generators split half (shared-memory port: 176 KB per step instead of 272).
constexpr uint32_t IDESC_TF32_PAIR = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (BN >> 3) << 17) | ((uint32_t) ((2 * BM) >> 4) << 24);
__device__ __forceinline__ void mma_tf32_pair(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t accumulate,
                                              uint32_t idesc = IDESC_TF32_PAIR) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit of the pair's MMAs, arriving on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void mma_commit_group2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t) 3)
                 : "memory");
}

// the same commit, arriving on the barrier at this offset in both CTAs of a 2-CTA cluster
__device__ __forceinline__ void mma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t) 3)
                 : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// x = hi + lo with hi the TF32 value nearest to x (ties away from zero) and lo the exact remainder
__device__ __forceinline__ void split_rn(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = __fsub_rn(x, hi);
}
// remainder after the truncation the tensor core applies to a raw fp32 operand
__device__ __forceinline__ float lo_trunc(float y) { return __fsub_rn(y, __uint_as_float(__float_as_uint(y) & 0xffffe000u)); }

// XMAT: the operator is materialised (S.buff, skge.hh:174-181 "buff != nullptr"): its 128 x 32 tiles come in by TMA
// (tmX, same K-major 128B-swizzled layout the generators write) on the stage's full barrier, and the generator warps
// only split them into the two TF32 terms.
// CL = 2: two CTAs that own neighbouring column tiles of C (same rows, same K range) form a cluster and share the
// generated operator tile: they take turns generating it (stage s is always produced by CTA s) and write it into BOTH
// shared memories (st.shared::cluster), arriving on both `ready` barriers; every MMA commit that frees a stage arrives on
// both `empty` barriers (multicast), so a stage is rewritten only when both tensor cores are done with it. Halves the
// generator work per flop (the binding resource for Gaussian operators).
// HALF: the two halves of the generator warps own one stage each (half h produces every K step with it % 2 == h: four
// Philox blocks per thread instead of two, and two step times to finish them), so tiles of two consecutive K steps are
// being generated at the same time. The generators are latency-bound, not issue-bound (see the cluster note above);
// this doubles the independent work in flight per scheduler. K-contiguous data, generated operator, no cluster.
// PAIR: cta_group::2. The two CTAs of a cluster (1, 2, 1) own two consecutive ROW tiles of C and the same column tile;
// CTA 0 (the leader) issues every MMA for both, each CTA generates its own X tile, loads and splits its half of Y,
// and signals the leader's `ready` barrier; the leader's commits free the stages of both CTAs (multicast).
template <bool GAUSS, bool XMAT, int CL = 1, bool HALF = false, bool PAIR = false>
__global__ void __launch_bounds__(TC_THREADS, 1) skge3_tc_kernel(const __grid_constant__ CUtensorMap tmY,
                                                                 const __grid_constant__ CUtensorMap tmX, const TcArgs a) {
    // stage geometry: a CTA of a pair holds half of the Y tile, which leaves room for a third stage in the same 192 KB
    constexpr int NST = PAIR ? 3 : STAGES;
    constexpr uint32_t YB = PAIR ? Y_BYTES / 2 : Y_BYTES;       // bytes of Y (and of Y_lo) per stage in this CTA
    constexpr uint32_t SB = 2 * X_BYTES + 2 * YB;               // bytes per stage
    static_assert(NST * SB <= RAW_OFFSET, "stages fit the dynamic shared memory the launcher asks for");
    static_assert(!(PAIR && HALF), "the two-halves generator schedule is tied to two stages");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(16) double2 logtab[GAUSS ? LOGF_TABLE_ENTRIES : 1];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t bar0 = base + BAR_OFFSET;
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_ready = [&](int s) { return bar0 + 8u * (NST + s); };
    auto bar_empty = [&](int s) { return bar0 + 8u * (2 * NST + s); };
    const uint32_t bar_accum = bar0 + 8u * (3 * NST);
    const uint32_t bar_raw_full = bar0 + 8u * (3 * NST + 1), bar_raw_empty = bar0 + 8u * (3 * NST + 2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFFSET + 8 * (3 * NST + 3));

    if constexpr (GAUSS) load_logf_table(logtab, a.logtab);
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bar_full(s), 1);
            // in a cluster the stage this CTA generates itself gets local arrivals only; the other one also the peer's
            uint32_t ready_count = HALF ? GEN_WARPS / 2 : GEN_WARPS;
            if constexpr (PAIR) {
                // leader: its generator warps + ONE arrival forwarded by the peer (its otherwise idle MMA warp waits on the
                // peer's local barrier and signals here: a cluster-scope release per generator warp and step cost a
                // membar each -- the top stall of the first version of this mode)
                uint32_t rk;
                asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rk));
                if (rk == 0) ready_count += 1;
            }
            if constexpr (CL > 1) {
                uint32_t rk;
                asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rk));
                if ((uint32_t) s != rk) ready_count = 2 * GEN_WARPS;
            }
            mbar_init(bar_ready(s), ready_count);
            mbar_init(bar_empty(s), PAIR ? 1 : CL);
        }
        mbar_init(bar_accum, 1);
        mbar_init(bar_raw_full, 1);
        mbar_init(bar_raw_empty, GEN_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
        if constexpr (XMAT) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    } else if (warp == 1) {
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    uint32_t crank = 0;
    if constexpr (CL > 1 || PAIR) {
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
        cluster_sync();                 // the peer's barriers exist before anything arrives on them
    }

    // the CTA pair of a cta_group::2 MMA is two CTAs adjacent in the cluster's (= the grid's) x dimension: in PAIR mode
    // the grid is (row tiles, column tiles, splits) instead of (column tiles, row tiles, splits)
    const int64_t i0 = (int64_t) (PAIR ? blockIdx.x : blockIdx.y) * BM, j0 = (int64_t) (PAIR ? blockIdx.y : blockIdx.x) * BN;
    const int split = blockIdx.z;
    const int per = a.steps_total / a.splits, rem = a.steps_total % a.splits;
    const int s_begin = split * per + min(split, rem);
    const int nsteps = per + (split < rem ? 1 : 0);

    if (warp == 0) {
        if (lane == 0) {
            constexpr int PF = 6;
            if (a.y_mn != 1) {
                const int jy = (int) j0 + (PAIR ? (int) crank * (BN / 2) : 0);      // first column this CTA loads
                constexpr int NBOX = (PAIR ? BN / 2 : BN) / 32;                     // y_mn == 2: boxes of 32 columns x 32 k
                // K-contiguous Y: one box of 32 k x (this CTA's columns); Q-contiguous Y (2): NBOX boxes, 4096 bytes apart
                auto prefetch_y = [&](int step) {
                    if (a.y_mn == 0) { tma_prefetch_2d(&tmY, step * BK, jy); return; }
                    for (int b = 0; b < NBOX; ++b) tma_prefetch_2d(&tmY, jy + 32 * b, step * BK);
                };
                auto load_y = [&](uint32_t dst, uint32_t bar, int step) {
                    if (a.y_mn == 0) { tma_load_2d(dst, &tmY, bar, step * BK, jy); return; }
                    for (int b = 0; b < NBOX; ++b) tma_load_2d(dst + 4096u * (uint32_t) b, &tmY, bar, jy + 32 * b, step * BK);
                };
                for (int it = 0; it < min(PF, nsteps); ++it) prefetch_y(s_begin + it);
                for (int it = 0; it < nsteps; ++it) {
                    const int st = it % NST;
                    const uint32_t ph = (it / NST) & 1;
                    if (it + PF < nsteps) prefetch_y(s_begin + it + PF);
                    mbar_wait(bar_empty(st), ph ^ 1);
                    // PAIR: this CTA's 128 columns of the tile and its own 128 rows of a materialised operator
                    mbar_arrive_expect_tx(bar_full(st), XMAT ? YB + X_BYTES : YB);
                    load_y(base + st * SB + 2 * X_BYTES, bar_full(st), s_begin + it);
                    if constexpr (XMAT) {
                        if (!a.x_mn)
                            tma_load_2d(base + st * SB, &tmX, bar_full(st), (int) a.xk0 + (s_begin + it) * BK, (int) (a.xr0 + i0));
                        else
                            for (int b = 0; b < BM / 32; ++b)      // row-contiguous operator: boxes of 32 rows x 32 k, 4096 bytes apart
                                tma_load_2d(base + st * SB + 4096u * (uint32_t) b, &tmX, bar_full(st), (int) (a.xr0 + i0) + 32 * b,
                                            (int) a.xk0 + (s_begin + it) * BK);
                    }
                }
            } else {
                // tensor map (Q, K), box 256 q (128 for a CTA of a pair) x RAW_K k, no swizzle: one raw tile in flight
                const int jy = (int) j0 + (PAIR ? (int) crank * (BN / 2) : 0);
                constexpr uint32_t RAWB = PAIR ? RAW_BYTES / 2 : RAW_BYTES;
                for (int j = 0; j < min(2 * PF, 2 * nsteps); ++j) tma_prefetch_2d(&tmY, jy, s_begin * BK + j * RAW_K);
                for (int j = 0; j < 2 * nsteps; ++j) {          // raw tile j = half (j & 1) of step j / 2
                    if (j + 2 * PF < 2 * nsteps) tma_prefetch_2d(&tmY, jy, s_begin * BK + (j + 2 * PF) * RAW_K);
                    if constexpr (XMAT) {
                        if ((j & 1) == 0) {                      // the X tile of step j / 2 rides that stage's full barrier
                            const int it = j >> 1, st = it % NST;
                            mbar_wait(bar_empty(st), (uint32_t) (((it / NST) & 1) ^ 1));
                            mbar_arrive_expect_tx(bar_full(st), X_BYTES);
                            tma_load_2d(base + st * SB, &tmX, bar_full(st), (int) a.xk0 + (s_begin + it) * BK,
                                        (int) (a.xr0 + i0));
                        }
                    }
                    mbar_wait(bar_raw_empty, (uint32_t) ((j & 1) ^ 1));
                    mbar_arrive_expect_tx(bar_raw_full, RAWB);
                    tma_load_2d(base + RAW_OFFSET, &tmY, bar_raw_full, jy, s_begin * BK + j * RAW_K);
                }
            }
        }
    } else if (warp == 1) {
        if constexpr (PAIR) {
            if (lane == 0 && crank == 0) {
                for (int it = 0; it < nsteps; ++it) {
                    const int st = it % NST;
                    const uint32_t ph = (it / NST) & 1;
                    // both CTAs' generator warps arrive here after their X tile, their half of Y (TMA, observed through
                    // their own full barrier) and its low part are in place
                    mbar_wait_cluster(bar_ready(st), ph);
                    if (XMAT || a.y_mn != 1) mbar_wait(bar_full(st), ph);
                    fence_proxy_async();
                    tc_fence_after();
                    const uint32_t xh = base + st * SB, xl = xh + X_BYTES, yh = xl + X_BYTES, yl = yh + YB;
                    const bool ymn = a.y_mn == 2;                 // Y tiles MN-major: 8 k rows = 1024 bytes per instruction
                    const bool xmn = XMAT && a.x_mn;              // row-contiguous materialised operator: X tiles MN-major
                    const uint32_t idesc = IDESC_TF32_PAIR | (ymn ? IDESC_B_MN : 0u) | (xmn ? IDESC_A_MN : 0u);
#pragma unroll
                    for (int kk = 0; kk < BK / 8; ++kk) {
                        const uint64_t dxh = xmn ? smem_desc_mn32(xh + kk * 1024) : smem_desc_sw128(xh + kk * 32);
                        const uint64_t dxl = xmn ? smem_desc_mn32(xl + kk * 1024) : smem_desc_sw128(xl + kk * 32);
                        const uint64_t dyh = ymn ? smem_desc_mn32(yh + kk * 1024) : smem_desc_sw128(yh + kk * 32);
                        const uint64_t dyl = ymn ? smem_desc_mn32(yl + kk * 1024) : smem_desc_sw128(yl + kk * 32);
                        const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
                        mma_tf32_pair(tmem_base + BN, dxl, dyh, acc, idesc);
                        mma_tf32_pair(tmem_base + BN, dxh, dyl, 1u, idesc);
                        mma_tf32_pair(tmem_base, dxh, dyh, acc, idesc);
                    }
                    mma_commit_group2(bar_empty(st));
                }
                mma_commit_group2(bar_accum);
            } else if (lane == 0) {
                // peer CTA: forward "my tiles of step it are ready" to the leader with one cluster-scope release
                const uint32_t leader_ready0 = map_to_cta(bar_ready(0), 0u);
                for (int it = 0; it < nsteps; ++it) {
                    const int st = it % NST;
                    mbar_wait(bar_ready(st), (uint32_t) ((it / NST) & 1));
                    mbar_arrive_remote(leader_ready0 + 8u * (uint32_t) st);
                }
            }
        } else
        if (lane == 0) {
            for (int it = 0; it < nsteps; ++it) {
                const int st = it % NST;
                const uint32_t ph = (it / NST) & 1;
                if constexpr (CL > 1) mbar_wait_cluster(bar_ready(st), ph);
                else mbar_wait(bar_ready(st), ph);
                if (XMAT || a.y_mn != 1) mbar_wait(bar_full(st), ph);
                if constexpr (CL > 1) fence_proxy_async();     // the peer's generic-proxy stores into this CTA's tiles
                tc_fence_after();
                const uint32_t xh = base + st * SB, xl = xh + X_BYTES, yh = xl + X_BYTES, yl = yh + YB;
                const bool ymn = a.y_mn == 2;
                const bool xmn = XMAT && a.x_mn;
                const uint32_t idesc = IDESC_TF32 | (ymn ? IDESC_B_MN : 0u) | (xmn ? IDESC_A_MN : 0u);
#pragma unroll
                for (int kk = 0; kk < BK / 8; ++kk) {
                    const uint64_t dxh = xmn ? smem_desc_mn32(xh + kk * 1024) : smem_desc_sw128(xh + kk * 32);
                    const uint64_t dxl = xmn ? smem_desc_mn32(xl + kk * 1024) : smem_desc_sw128(xl + kk * 32);
                    const uint64_t dyh = ymn ? smem_desc_mn32(yh + kk * 1024) : smem_desc_sw128(yh + kk * 32);
                    const uint64_t dyl = ymn ? smem_desc_mn32(yl + kk * 1024) : smem_desc_sw128(yl + kk * 32);
                    const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
                    mma_tf32(tmem_base + BN, dxl, dyh, acc, idesc);
                    mma_tf32(tmem_base + BN, dxh, dyl, 1u, idesc);
                    mma_tf32(tmem_base, dxh, dyh, acc, idesc);
                }
                if constexpr (CL > 1) mma_commit_pair(bar_empty(st));
                else mma_commit(bar_empty(st));
            }
            mma_commit(bar_accum);
        }
    } else {
        // ---------------- generators ----------------
        const int gt = threadIdx.x - 64;
        const uint64_t seed_lo = ((uint64_t) a.ctr.c1 << 32) | a.ctr.c0, seed_hi = ((uint64_t) a.ctr.c3 << 32) | a.ctr.c2;
        if constexpr (HALF) {
            constexpr int HT = 16 * GEN_WARPS;                 // threads per half
            constexpr int HB = (BM * BK / 4) / HT;             // Philox blocks per thread and tile
            const int half = gt / HT, ht = gt % HT;
            const int hc = ht & 7, hr = ht >> 3;               // 16-byte chunk of the row, first row; rows hr + (HT / 8) * rr
            uint64_t hoff[HB];
#pragma unroll
            for (int rr = 0; rr < HB; ++rr)
                hoff[rr] = (uint64_t) ((a.v0 + i0 + hr + (HT / 8) * rr) * a.R + a.ublk0 + hc) + 8ull * (uint64_t) (s_begin + half);
            const uint32_t hxoff = (uint32_t) hr * 128u + (uint32_t) ((hc ^ (hr & 7)) << 4);
            uint8_t* stage = smem + half * SB;        // stage == half, always
            for (int it = half; it < nsteps; it += 2) {
                const uint32_t ph = (uint32_t) ((it >> 1) & 1);
                mbar_wait(bar_empty(half), ph ^ 1);
#pragma unroll
                for (int rr = 0; rr < HB; ++rr) {
                    const uint64_t lo = seed_lo + hoff[rr];
                    const uint64_t hi = seed_hi + (lo < seed_lo ? 1ull : 0ull);
                    hoff[rr] += 16;
                    const Ctr128 cc{(uint32_t) lo, (uint32_t) (lo >> 32), (uint32_t) hi, (uint32_t) (hi >> 32)};
                    float4 f = transform4<GAUSS>(philox4x32_10(cc, a.key), logtab);
                    if (a.kshift) {
                        const float4 g = transform4<GAUSS>(philox4x32_10(ctr_add(cc, 1), a.key), logtab);
                        if (a.kshift == 1) f = make_float4(f.y, f.z, f.w, g.x);
                        else if (a.kshift == 2) f = make_float4(f.z, f.w, g.x, g.y);
                        else f = make_float4(f.w, g.x, g.y, g.z);
                    }
                    float4 h, l;
                    split_rn(finish_sample<float, GAUSS>(f.x), h.x, l.x);
                    split_rn(finish_sample<float, GAUSS>(f.y), h.y, l.y);
                    split_rn(finish_sample<float, GAUSS>(f.z), h.z, l.z);
                    split_rn(finish_sample<float, GAUSS>(f.w), h.w, l.w);
                    *reinterpret_cast<float4*>(stage + hxoff + rr * ((HT / 8) * 128)) = h;
                    *reinterpret_cast<float4*>(stage + X_BYTES + hxoff + rr * ((HT / 8) * 128)) = l;
                }
                mbar_wait(bar_full(half), ph);
                const uint8_t* ysrc = stage + 2 * X_BYTES;
                uint8_t* ydst = stage + 2 * X_BYTES + YB;
#pragma unroll
                for (int q = 0; q < (int) (YB / 16) / HT; ++q) {
                    const uint32_t o = (uint32_t) (ht + HT * q) * 16u;
                    const float4 y = *reinterpret_cast<const float4*>(ysrc + o);
                    float4 l;
                    l.x = lo_trunc(y.x); l.y = lo_trunc(y.y); l.z = lo_trunc(y.z); l.w = lo_trunc(y.w);
                    *reinterpret_cast<float4*>(ydst + o) = l;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_ready(half));
            }
        } else {
        const int c = gt & 7, r0 = gt >> 3;
        constexpr int ROWS_PER_PASS = 4 * GEN_WARPS;      // rows of X covered by one pass of all generator threads
        // In a cluster the two CTAs take turns: CTA r generates the whole tile of every K step whose stage is r and writes
        // it into both shared memories. (Splitting the ROWS of each tile between the CTAs was measured first and did not
        // help: the generators are latency-bound -- 4 warps per scheduler, two dependent Box-Muller chains per thread --
        // so one block per thread takes as long as two. Alternating keeps two blocks per thread and gives each CTA two
        // step times to produce a tile.)
        constexpr int XPT = X_PER_THREAD;
        const int rbase = r0;
        uint64_t off[XPT];
#pragma unroll
        for (int rr = 0; rr < XPT; ++rr)
            off[rr] = (uint64_t) ((a.v0 + i0 + rbase + ROWS_PER_PASS * rr) * a.R + a.ublk0 + c) + 8ull * (uint64_t) s_begin;
        const uint32_t xoff = (uint32_t) rbase * 128u + (uint32_t) ((c ^ (r0 & 7)) << 4);
        for (int it = 0; it < nsteps; ++it) {
            const int st = it % NST;
            const uint32_t ph = (it / NST) & 1;
            uint8_t* stage = smem + st * SB;
            mbar_wait(bar_empty(st), ph ^ 1);
            if (a.y_mn == 1) {
                // transpose the raw Q-contiguous tile into the K-major swizzled Y (the tensor core truncates it to TF32
                // itself) and Y_lo tiles of this stage, then hand the raw buffer back to the TMA producer
                const uint8_t* rawt = smem + RAW_OFFSET;
                uint8_t* yh = stage + 2 * X_BYTES;
                uint8_t* yl = yh + YB;
#pragma unroll
                for (int half = 0; half < BK / RAW_K; ++half) {
                    mbar_wait(bar_raw_full, (uint32_t) half);          // raw tile 2 it + half: parity = half
                    constexpr int BNL = PAIR ? BN / 2 : BN;               // columns of the tile held by this CTA
#pragma unroll
                    for (int q4 = 0; q4 < (BNL * RAW_K / 4) / (32 * GEN_WARPS); ++q4) {
                        const int ch = gt + 32 * GEN_WARPS * q4;
                        const int qq = ch & (BNL - 1), kl = ch / BNL;      // column of the tile, 4-deep k chunk of this half
                        const float* src = reinterpret_cast<const float*>(rawt) + (4 * kl) * BNL + qq;
                        float4 y;
                        y.x = src[0]; y.y = src[BNL]; y.z = src[2 * BNL]; y.w = src[3 * BNL];
                        float4 l;
                        l.x = lo_trunc(y.x); l.y = lo_trunc(y.y); l.z = lo_trunc(y.z); l.w = lo_trunc(y.w);
                        const int kc = kl + half * (RAW_K / 4);
                        const uint32_t o = (uint32_t) qq * 128u + (uint32_t) ((kc ^ (qq & 7)) << 4);
                        *reinterpret_cast<float4*>(yh + o) = y;
                        *reinterpret_cast<float4*>(yl + o) = l;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_raw_empty);
                }
            }
            if constexpr (XMAT) {
                // materialised operator: split the raw tile TMA delivered into hi (nearest TF32, rewritten in place) and lo
                mbar_wait(bar_full(st), ph);
#pragma unroll
                for (int q = 0; q < (int) (X_BYTES / 16) / (32 * GEN_WARPS); ++q) {
                    const uint32_t o = (uint32_t) (gt + 32 * GEN_WARPS * q) * 16u;
                    const float4 x = *reinterpret_cast<const float4*>(stage + o);
                    float4 h, l;
                    split_rn(x.x, h.x, l.x); split_rn(x.y, h.y, l.y); split_rn(x.z, h.z, l.z); split_rn(x.w, h.w, l.w);
                    *reinterpret_cast<float4*>(stage + o) = h;
                    *reinterpret_cast<float4*>(stage + X_BYTES + o) = l;
                }
            } else {
            const bool my_turn = (CL == 1) || ((uint32_t) st == crank);
            if (a.x_t) {
                // Philox blocks along the rows of X: thread -> (column k of the step, 4-row block); element (i, k) is lane
                // (u0 + i0 + i) & 3 of block (v0 + k) * R + (u0 + i0 + i) / 4 with u0 % 4 == 0 (checked by the launcher).
                // The K-major swizzled tile is written with 4-byte stores: the 32 lanes of a warp hold 32 consecutive k
                // of the same rows, i.e. 8 swizzled 16-byte chunks x 4 words = 32 different banks per store.
                const int kl = gt & 31, rb0 = gt >> 5;
#pragma unroll
                for (int rr = 0; rr < XPT; ++rr) {
                    const int rb = rb0 + (GEN_WARPS) * rr;
                    const uint64_t o = (uint64_t) ((a.v0 + (int64_t) (s_begin + it) * BK + kl) * a.R + a.ublk0 + (i0 >> 2) + rb);
                    const uint64_t lo = seed_lo + o;
                    const uint64_t hi = seed_hi + (lo < seed_lo ? 1ull : 0ull);
                    const Ctr128 cc{(uint32_t) lo, (uint32_t) (lo >> 32), (uint32_t) hi, (uint32_t) (hi >> 32)};
                    const float4 f = transform4<GAUSS>(philox4x32_10(cc, a.key), logtab);
                    const float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int row = 4 * rb + j;
                        float h, l;
                        split_rn(finish_sample<float, GAUSS>(fv[j]), h, l);
                        const uint32_t o4 = (uint32_t) row * 128u + (uint32_t) (((kl >> 2) ^ (row & 7)) << 4) + (uint32_t) (kl & 3) * 4u;
                        *reinterpret_cast<float*>(stage + o4) = h;
                        *reinterpret_cast<float*>(stage + X_BYTES + o4) = l;
                    }
                }
            } else {
#pragma unroll
            for (int rr = 0; rr < XPT; ++rr) {
                const uint64_t lo = seed_lo + off[rr];
                const uint64_t hi = seed_hi + (lo < seed_lo ? 1ull : 0ull);
                off[rr] += 8;
                if (!my_turn) continue;
                const Ctr128 cc{(uint32_t) lo, (uint32_t) (lo >> 32), (uint32_t) hi, (uint32_t) (hi >> 32)};
                float4 f = transform4<GAUSS>(philox4x32_10(cc, a.key), logtab);
                if (a.kshift) {
                    // window origin not on a Philox block boundary: the four positions straddle two blocks
                    const float4 g = transform4<GAUSS>(philox4x32_10(ctr_add(cc, 1), a.key), logtab);
                    if (a.kshift == 1) f = make_float4(f.y, f.z, f.w, g.x);
                    else if (a.kshift == 2) f = make_float4(f.z, f.w, g.x, g.y);
                    else f = make_float4(f.w, g.x, g.y, g.z);
                }
                float4 h, l;
                split_rn(finish_sample<float, GAUSS>(f.x), h.x, l.x);
                split_rn(finish_sample<float, GAUSS>(f.y), h.y, l.y);
                split_rn(finish_sample<float, GAUSS>(f.z), h.z, l.z);
                split_rn(finish_sample<float, GAUSS>(f.w), h.w, l.w);
                *reinterpret_cast<float4*>(stage + xoff + rr * (ROWS_PER_PASS * 128)) = h;
                *reinterpret_cast<float4*>(stage + X_BYTES + xoff + rr * (ROWS_PER_PASS * 128)) = l;
                if constexpr (CL > 1) {
                    const uint32_t ra = map_to_cta(base + (uint32_t) st * SB + xoff + rr * (ROWS_PER_PASS * 128), crank ^ 1u);
                    st_cluster_v4(ra, h);
                    st_cluster_v4(ra + X_BYTES, l);
                }
            }
            }   // !x_t
            }
            if (a.y_mn != 1) {          // the low part of Y element by element: the same code for K-major and MN-major tiles
            mbar_wait(bar_full(st), ph);
            const uint8_t* ysrc = stage + 2 * X_BYTES;
            uint8_t* ydst = stage + 2 * X_BYTES + YB;
#pragma unroll
            for (int q = 0; q < (int) (YB / 16) / (32 * GEN_WARPS); ++q) {
                const uint32_t o = (uint32_t) (gt + 32 * GEN_WARPS * q) * 16u;
                const float4 y = *reinterpret_cast<const float4*>(ysrc + o);
                float4 l;
                l.x = lo_trunc(y.x); l.y = lo_trunc(y.y); l.z = lo_trunc(y.z); l.w = lo_trunc(y.w);
                *reinterpret_cast<float4*>(ydst + o) = l;
            }
            }
            // PAIR: the tiles were written into THIS CTA's shared memory (the pair's tensor cores read them through the async
            // proxy), so the shared::cta fence is enough; the all-state-space fence of the CL = 2 mode costs a full membar
            if constexpr (CL > 1) fence_proxy_async_all();
            else fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_ready(st));
                if constexpr (CL > 1) {
                    if ((uint32_t) st == crank) mbar_arrive_remote(map_to_cta(bar_ready(st), crank ^ 1u));
                }
            }
        }
        }   // !HALF
        // ---------------- epilogue ----------------
        mbar_wait(bar_accum, 0);
        tc_fence_after();
        constexpr int COLS_PER_WARP = BN / (GEN_WARPS / 4);
        const int q4 = warp & 3, part = (warp - 2) >> 2;
        const int64_t i = i0 + q4 * 32 + lane;
        for (int cb = 0; cb < COLS_PER_WARP / 32; ++cb) {
            const int col0 = part * COLS_PER_WARP + cb * 32;
            uint32_t v[32];
            {
                uint32_t vs[32];
                tmem_ld32(tmem_base + ((uint32_t) (q4 * 32) << 16) + (uint32_t) col0, v);
                tmem_ld32(tmem_base + ((uint32_t) (q4 * 32) << 16) + (uint32_t) (BN + col0), vs);
#pragma unroll
                for (int t = 0; t < 32; ++t) v[t] = __float_as_uint(__fadd_rn(__uint_as_float(v[t]), __uint_as_float(vs[t])));
            }
            if (a.W) {
                float* w = a.W + ((int64_t) split * a.Q_pad + j0 + col0) * a.P_pad + i;
#pragma unroll
                for (int t = 0; t < 32; ++t) w[(int64_t) t * a.P_pad] = __uint_as_float(v[t]);
            } else if (i < a.P) {
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const int64_t j = j0 + col0 + t;
                    if (j < a.Q) {
                        float* cp = a.C + i * a.crs + j * a.ccs;
                        float r = a.alpha * __uint_as_float(v[t]);
                        if (a.beta != 0.f) r += a.beta * (*cp);
                        *cp = r;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CL > 1 || PAIR) cluster_sync();   // no CTA leaves while its peer can still arrive on its barriers
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// C = alpha * sum_s W[s] + beta * C, summed in split order (deterministic)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ W, int splits, int64_t P, int64_t Q,
                                                            int64_t P_pad, int64_t Q_pad, float alpha, float beta,
                                                            float* __restrict__ C, int64_t crs, int64_t ccs) {
    const int64_t total = P * Q;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
        const int64_t j = e / P, i = e - j * P;
        const float* w = W + j * P_pad + i;
        float s = 0.f;
#pragma unroll 8
        for (int sp = 0; sp < splits; ++sp) s += w[(int64_t) sp * Q_pad * P_pad];
        float* cp = C + i * crs + j * ccs;
        float r = alpha * s;
        if (beta != 0.f) r += beta * (*cp);
        *cp = r;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn) p;
        else
            cudaGetLastError();
    }
    return fn;
}

}  // namespace

// Gaussian operators with several column tiles: every column tile of C regenerates the same 128 x 32 operator tile, and
// the Box-Muller generator (not the tensor core) bounds the fused kernel (d = n = 1024, m = 1e5: 1.70 ms against 0.97 ms for
// Uniform). With two or more column tiles it is cheaper to generate each K panel of op(S) ONCE into a scratch buffer with the fill
// kernel and run the materialised-operator (XMAT) instantiation on it: same operand values and, for a single panel, the
// same split-K as the fused kernel. The panel (<= 512 MB of scratch) is written and re-read immediately and never becomes a
// caller-visible S. "tc_materialise": 0 never, 1 (default) auto, 2 always.
static int dense_tc_f32_via_panel(const DenseProblem<float>& p, bool x_t, cudaStream_t st) {
    int64_t kp = ((int64_t) 512 << 20) / 4 / (p.P > 0 ? p.P : 1);
    kp = (kp / 1280) * 1280;                        // whole 40-step chains of 32
    if (kp < 1280) kp = 1280;
    if (kp > p.K) kp = (p.K + 3) / 4 * 4;
    const int64_t ld = kp;
    float* panel = (float*) workspace(5, (size_t) p.P * (size_t) ld * sizeof(float), st);
    if (!panel) return -2;                          // no room for the scratch panel: the caller runs the fused kernel
    for (int64_t k0 = 0; k0 < p.K; k0 += kp) {
        const int64_t kc = (p.K - k0 < kp) ? p.K - k0 : kp;
        int rc;
        if (!x_t) rc = launch_fill_dense<float>(p.gen, p.family, p.v0, p.P, p.u0 + k0, kc, panel, ld, 1, st);
        else rc = launch_fill_dense<float>(p.gen, p.family, p.v0 + k0, kc, p.u0, p.P, panel, 1, ld, st);
        if (rc) return rc;
        DenseProblem<float> q = p;
        q.S_buff = panel; q.S_ld = ld;
        q.v0 = 0; q.u0 = 0; q.vi = 1; q.ui = 0; q.vk = 0; q.uk = 1;
        q.K = kc;
        q.Y = p.Y + k0 * p.yrs;
        q.beta = (k0 == 0) ? p.beta : 1.0f;
        rc = launch_dense_tc_f32(q, st);
        if (rc) return rc < 0 ? fail("operator panel: the tensor-core kernel refused its own panel") : rc;
    }
    return 0;
}

int launch_dense_tc_f32(const DenseProblem<float>& p, cudaStream_t st) {
    // shapes / layouts this kernel takes; everything else goes to the generic kernel
    const bool xmat = p.S_buff != nullptr;
    if (xmat && get_option("dense_path") == 4) return -1;     // experiment switch: materialised operators to the generic kernel
    if (!xmat && p.family == 'G' && !p.gen.logtab) return -1;
    // Philox blocks along K (the default Axis::Long operators used as they are), or along the rows of X (Axis::Short
    // operators and transposed uses: reference case table dense_skops.hh:187-199) when the window starts on a block
    // A materialised operator whose vectors run along the rows of X (a filled Axis::Short operator) is read as an MN-major A operand.
    const bool x_t = (p.ui == 1 && p.vk == 1);
    if (!(p.uk == 1 && p.vi == 1) && !(x_t && (p.u0 & 3) == 0)) return -1;
    if (x_t && xmat && get_option("tc_xmn") == 0) return -1;  // experiment switch: row-contiguous materialised operators to the generic kernel
    // Y K-contiguous, or Q-contiguous (left sketch of RowMajor data, right sketch of ColMajor data)
    const bool y_mn = (p.yrs != 1);
    if (y_mn && p.ycs != 1) return -1;
    if (y_mn && get_option("dense_path") == 3) return -1;     // experiment switch: Q-contiguous data to the generic kernel
    if (p.K < 64 || p.P < 1 || p.Q < 1) return -1;
    if (p.Q > 0x7fffff00LL || p.K > 0x7fffff00LL) return -1;
    if ((reinterpret_cast<uintptr_t>(p.Y) & 15) != 0 || ((y_mn ? p.yrs : p.ycs) & 3) != 0) return -1;   // TMA alignment rules
    if ((int64_t) p.P * p.Q < 128 * 64 && p.K < 4096) return -1;                       // tiny: launch cost dominates
    if (xmat) {
        // vector v of the operator is the contiguous run S_buff[v * S_ld ...]: X(i, k) = S_buff[(v0 + i) * S_ld + u0 + k]
        // TMA: base, row pitch and the box's first element must all sit on 16-byte boundaries (a window whose first
        // column is not a multiple of 4 trapped the pipeline on the device), so such windows go to the generic kernel
        // (row-contiguous operators, x_t: X(i, k) = S_buff[(v0 + k) * S_ld + u0 + i], the same rules with the roles of i and k swapped)
        if ((reinterpret_cast<uintptr_t>(p.S_buff) & 15) != 0 || (p.S_ld & 3) != 0 || (p.u0 & 3) != 0) return -1;
        if (p.v0 + (x_t ? p.K : p.P) > 0x7fffff00LL || p.u0 + (x_t ? p.P : p.K) > 0x7fffff00LL) return -1;
    }
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return -1;

    const int64_t tiles_p = (p.P + BM - 1) / BM, tiles_q = (p.Q + BN - 1) / BN;
    if (tiles_p > 65535 || tiles_q > 0x7fffffff) return -1;
    {
        const int64_t mat_opt = get_option("tc_materialise");
        if (!xmat && p.family == 'G' && (mat_opt == 2 || (mat_opt == 1 && tiles_q >= 2)) && p.K >= 64) {
            const int prc = dense_tc_f32_via_panel(p, x_t, st);
            if (prc != -2) return prc;                  // -2: no memory for the panel, run fused
        }
    }
    const int kshift = x_t ? 0 : (int) (p.u0 & 3);
    const int64_t steps = (p.K + BK - 1) / BK;
    // Pairs of column tiles share the generated operator tile through a 2-CTA cluster. tc_cluster: 0 never; 1 (default)
    // where it was measured to pay -- Gaussian operators with K-contiguous data (1.83 -> 1.70 ms at d = n = 1024,
    // m = 1e5); 2 whenever the tile count is even. It does not pay elsewhere: distributed shared memory moves ~21 B/clk,
    // so shipping the 32 KB (hi, lo) tile costs as much as the MMAs of the step (Uniform: 1.06 -> 1.44 ms).
    const int64_t cl_opt = get_option("tc_cluster");
    // CTA pairs (cta_group::2, "tc_pair", default on): two consecutive row tiles share the Y tile. Generated operator,
    // K-contiguous data, an even number of row tiles.
    const int64_t pair_opt = get_option("tc_pair");
    // tc_pair: 0 never, 1 (default) where it was measured to pay -- Uniform operators (C1: 1.05 -> 0.98 ms; Gaussian operators
    // are bound by the generator warps and lose the two-halves / shared-tile schedules: 1.79 -> 1.94 ms), 2 whenever possible
    static thread_local bool pair_refused = false;     // a refused cluster launch (e.g. a partitioned GPU) falls back once and for all
    const bool pair = !pair_refused && (tiles_p % 2 == 0) && tiles_q <= 65535 && cl_opt != 2 &&
                      (pair_opt == 2 || (pair_opt == 1 && (xmat || p.family == 'U')));
    const bool cluster = !pair && !xmat && !x_t && (tiles_q % 2 == 0) && (cl_opt == 2 || (cl_opt == 1 && p.family == 'G' && !y_mn));
    // Split K. Two constraints: (1) the tensor core adds into its fp32 accumulator with truncation, a bias that
    // grows linearly with the number of accumulated MMAs (measured: 4.6e-4 relative after 2048 K steps, 1.9e-6
    // after 8), so no partial sum stays in TMEM for more than MAX_CHAIN_STEPS steps; partial sums are added in
    // round-to-nearest fp32 by the reduce kernel. (2) the grid should cover the SMs in whole waves.
    const int64_t tiles = tiles_p * tiles_q;
    const int sms = sm_count();
    int64_t s_min = (steps + MAX_CHAIN_STEPS - 1) / MAX_CHAIN_STEPS;
    int splits = (int) s_min;
    {
        int64_t ctas = tiles * s_min;
        int64_t waves = (ctas + sms - 1) / sms;
        int64_t s_fill = waves * sms / tiles;          // more splits that still fit the same number of waves
        if (s_fill > steps / 4) s_fill = steps / 4;    // keep at least 4 steps per CTA
        if (s_fill > s_min) splits = (int) s_fill;
        if (splits < 1) splits = 1;
    }
    if (get_option("tc_splits") > 0) {
        splits = (int) get_option("tc_splits");
        if (splits > steps) splits = (int) steps;
    }
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t) (y_mn ? p.Q : p.K), (cuuint64_t) (y_mn ? p.K : p.Q)};
    const cuuint64_t gstr[1] = {(cuuint64_t) (y_mn ? p.yrs : p.ycs) * 4ull};
    // Q-contiguous data: "tc_ymn" 0 (default) = tiles go to the tensor core as they are, as an MN-major operand (boxes of 32 q x
    // 32 k, 128B swizzle with 32-byte atoms); 1 = the transposing path of the first half of round 2 (raw boxes of 256 q x 16 k)
    const int ymode = !y_mn ? 0 : ((get_option("tc_ymn") == 1 && !(xmat && x_t)) ? 1 : 2);
    const cuuint32_t box[2] = {(cuuint32_t) (ymode == 2 ? 32 : ymode == 1 ? (pair ? BN / 2 : BN) : BK),
                               (cuuint32_t) (ymode == 2 ? BK : ymode == 1 ? RAW_K : (pair ? BN / 2 : BN))};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.Y), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE,
                      ymode == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : ymode == 1 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return -1;

    CUtensorMap tmx = tm;
    if (xmat) {
        // extents end at the window's last column / row so that the K tail and the rows past P read as zeros
        const cuuint64_t xdim[2] = {(cuuint64_t) (p.u0 + (x_t ? p.P : p.K)), (cuuint64_t) (p.v0 + (x_t ? p.K : p.P))};
        const cuuint64_t xstr[1] = {(cuuint64_t) p.S_ld * 4ull};
        // K-contiguous operator: one box of 32 k x 128 rows, 128B swizzle; row-contiguous: boxes of 32 rows x 32 k, 32-byte atoms
        const cuuint32_t xbox[2] = {(cuuint32_t) (x_t ? 32 : BK), (cuuint32_t) (x_t ? BK : BM)};
        cr = enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.S_buff), xdim, xstr, xbox, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, x_t ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return -1;
    }

    TcArgs a;
    a.xr0 = (xmat && x_t) ? p.u0 : p.v0; a.xk0 = (xmat && x_t) ? p.v0 : p.u0;
    a.x_mn = (xmat && x_t) ? 1 : 0;
    a.ctr = p.gen.ctr; a.key = p.gen.key; a.R = p.gen.R; a.logtab = p.gen.logtab;
    a.v0 = p.v0;
    a.kshift = kshift;
    a.ublk0 = p.u0 >> 2;
    a.P = p.P; a.Q = p.Q;
    a.steps_total = (int) steps;
    a.y_mn = ymode;
    a.x_t = x_t ? 1 : 0;
    a.splits = splits;
    a.alpha = p.alpha; a.beta = p.beta;
    a.C = p.C; a.crs = p.crs; a.ccs = p.ccs;
    a.P_pad = tiles_p * BM; a.Q_pad = tiles_q * BN;
    a.W = nullptr;
    if (splits > 1) {
        a.W = (float*) workspace(6, (size_t) splits * a.P_pad * a.Q_pad * sizeof(float), st);
        if (!a.W) return fail_cuda(cudaErrorMemoryAllocation, "split-K workspace");
    }
    static DevOnce attr_done[10];
    const bool gauss = p.family == 'G';
    // tc_halves: 1 (default) = the generator warps work on two K steps at a time where the kernel supports it
    const bool halves = !xmat && !cluster && !pair && !y_mn && !x_t && get_option("tc_halves") != 0;
    const int variant = xmat ? (pair ? 9 : 2) : (pair ? 7 : (cluster ? 3 : (halves ? 5 : 0))) + (gauss ? 1 : 0);
    auto set_attr = [&](auto kern) { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM); };
    if (attr_done[variant].need()) {
        cudaError_t e;
        switch (variant) {
            case 0: e = set_attr(skge3_tc_kernel<false, false, 1, false>); break;
            case 1: e = set_attr(skge3_tc_kernel<true, false, 1, false>); break;
            case 2: e = set_attr(skge3_tc_kernel<false, true, 1, false>); break;
            case 3: e = set_attr(skge3_tc_kernel<false, false, 2, false>); break;
            case 4: e = set_attr(skge3_tc_kernel<true, false, 2, false>); break;
            case 5: e = set_attr(skge3_tc_kernel<false, false, 1, true>); break;
            case 6: e = set_attr(skge3_tc_kernel<true, false, 1, true>); break;
            case 7: e = set_attr(skge3_tc_kernel<false, false, 1, false, true>); break;
            case 8: e = set_attr(skge3_tc_kernel<true, false, 1, false, true>); break;
            default: e = set_attr(skge3_tc_kernel<false, true, 1, false, true>); break;
        }
        if (e != cudaSuccess) { cudaGetLastError(); return -1; }
        attr_done[variant].done();
    }
    dim3 grid((unsigned) tiles_q, (unsigned) tiles_p, (unsigned) splits);
    if (pair) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned) tiles_p, (unsigned) tiles_q, (unsigned) splits);
        cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TC_SMEM; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e;
        if (xmat) e = cudaLaunchKernelEx(&cfg, skge3_tc_kernel<false, true, 1, false, true>, tm, tmx, a);
        else e = gauss ? cudaLaunchKernelEx(&cfg, skge3_tc_kernel<true, false, 1, false, true>, tm, tmx, a)
                       : cudaLaunchKernelEx(&cfg, skge3_tc_kernel<false, false, 1, false, true>, tm, tmx, a);
        if (e != cudaSuccess) {
            cudaGetLastError();
            pair_refused = true;
            return launch_dense_tc_f32(p, st);          // same problem, single-CTA kernels
        }
    } else
    if (xmat) skge3_tc_kernel<false, true, 1, false><<<grid, TC_THREADS, TC_SMEM, st>>>(tm, tmx, a);
    else if (halves) {
        if (gauss) skge3_tc_kernel<true, false, 1, true><<<grid, TC_THREADS, TC_SMEM, st>>>(tm, tmx, a);
        else skge3_tc_kernel<false, false, 1, true><<<grid, TC_THREADS, TC_SMEM, st>>>(tm, tmx, a);
    } else if (!cluster) {
        if (gauss) skge3_tc_kernel<true, false, 1, false><<<grid, TC_THREADS, TC_SMEM, st>>>(tm, tmx, a);
        else skge3_tc_kernel<false, false, 1, false><<<grid, TC_THREADS, TC_SMEM, st>>>(tm, tmx, a);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TC_SMEM; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = gauss ? cudaLaunchKernelEx(&cfg, skge3_tc_kernel<true, false, 2, false>, tm, tmx, a)
                              : cudaLaunchKernelEx(&cfg, skge3_tc_kernel<false, false, 2, false>, tm, tmx, a);
        if (e != cudaSuccess) return fail_cuda(e, "cluster launch of the tensor-core sketch kernel");
    }
    count_launch();
    count_tc_launch();
    RB_CUDA(cudaGetLastError());
    if (splits > 1) {
        int64_t g = (p.P * p.Q + 255) / 256;
        if (g > (int64_t) sms * 8) g = (int64_t) sms * 8;
        splitk_reduce_kernel<<<(unsigned) g, 256, 0, st>>>(a.W, splits, p.P, p.Q, a.P_pad, a.Q_pad, p.alpha, p.beta, p.C,
                                                          p.crs, p.ccs);
        count_launch();
        RB_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace rb
