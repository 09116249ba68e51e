// K2b: fused generate-and-multiply dense sketch for double on the FP64 tensor-core path (DMMA).
//
// Replaces dense::lskge3 / rskge3 (RandBLAS/skge.hh:154-202, 307-355) for double when the operator is not
// materialised (the reference fills a d x m host buffer -- 131 GB at BASELINE config 3 -- and calls blas::gemm).
// tcgen05.mma has no f64 kind (ptxas rejects kind::f64 for sm_100a), so the FP64 tensor path on Blackwell is the
// warp-level mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) with register accumulators; DMMA is an IEEE fused
// multiply-add chain, so unlike the TF32 path there is no accumulator-truncation issue and split-K is only used
// to fill the SMs.
//
// Canonical problem (kernels.h): C(P x Q) = alpha * X(P x K) * Y(K x Q) + beta * C, X = op(S window).
// CTA tile 64 x 256 (Q > 128) or 128 x 128, 8 warps, warp tile 64 x 32 = 8 x 4 DMMA tiles (64 accumulator doubles
// per thread), K step 16 (see DmmaTile below for why the wide tile). 16 warps with 32 x 32 warp tiles measured the
// same as 8: the DMMA pipe does not get busier with four warps per scheduler than with two. Per step every thread
//   * issues the 16-byte cp.async copies of the next Y tile (columns of A are K-contiguous: ColMajor A),
//   * runs the 128 DMMAs of the current tile from shared memory (rows padded to 20 doubles: conflict-free
//     8-byte fragment loads),
//   * regenerates its share (1 or 2 Philox blocks) of the next DM x 16 tile of S from
//     (key, counter, ro_s, co_s) -- Philox4x32-10, uneg11 or Box-Muller in float exactly as fill_dense, promoted
//     to double; S never touches HBM.
// Three shared-memory stages and mbarriers (full[]: one arrival per thread + one per thread's cp.async completions,
// empty[]: one arrival per warp) instead of a __syncthreads per step, and the two warps that share a scheduler run OUT OF PHASE: warps 0-3
// do [DMMAs of step s][generate step s+2], warps 4-7 do [generate step s+2][DMMAs of step s] (3% faster). A warp issues in
// order, so while it generates (~350 instructions per step) it feeds no DMMAs; with the lockstep version both warps
// of a scheduler generated at the same time and the FP64 tensor pipe idled for that part of every step.
// Measured and rejected: putting the generation arithmetic INTO the DMMA sequence of the same warp (stages tied to
// accumulator values so that ptxas cannot hoist them; the DMMA block then carries the Philox / Box-Muller instructions
// instead of NOPs): 84.3 ms vs 73.4 ms at C3's shape -- DMMAs want to issue back to back, the other warp of the
// scheduler is the better filler.
//
// Roofline: FP64 tensor. 2*P*Q*K flops per launch.
#include "common.cuh"
#include "kernels.h"
#include "tma.cuh"

namespace rb {

namespace {

constexpr int DK = 16, DLD = DK + 4;                         // K step; padded row length in doubles. DK = 32 with two
                                                              // stages: Uniform operator 70.0 -> 66.2 ms at C3's shape, Gaussian 72.9 -> 73.7
constexpr int D_CPR = DK / 4;                                 // 4-wide chunks (Philox blocks) per row of a tile
constexpr int D_THREADS = 256;
constexpr int D_STAGES = 3;
constexpr int D_AHEAD = D_STAGES - 1;                         // a warp produces stage s + D_AHEAD in the iteration that consumes s
// Two CTA tiles, both with 8 warps of 64 x 32 (8 x 4 DMMA tiles, 64 accumulator doubles per thread):
//   64 x 256 (warps 1 x 8): half the operator rows per column of C, i.e. half the generation work per flop -- the
//             Gaussian operator costs 13% of the kernel at 128 x 128 and 4% here (C3: 81.3 -> 72.9 ms);
//   128 x 128 (warps 2 x 4): for Q <= 128, where the wide tile would be half empty.
template <int DM_, int DN_, int WM_, int WN_>
struct DmmaTile {
    static constexpr int DM = DM_, DN = DN_, D_WM = WM_, D_WN = WN_;
    static constexpr int D_MI = DM / (D_WM * 8), D_NI = DN / (D_WN * 8);
    static constexpr int D_GR = DM / (D_THREADS / D_CPR);    // rows of the S tile generated per thread and step
    static constexpr int DLQ = DN + 4;                       // row length of a Q-contiguous Y tile [k][q] (conflict-free fragments)
    static constexpr size_t smem = (size_t) D_STAGES * (DM + DN) * DLD * sizeof(double);
    static_assert(D_WM * D_WN * 32 == D_THREADS && D_GR >= 1 && DK * DLQ <= DN * DLD, "tile configuration");
};
using TileWide = DmmaTile<64, 256, 1, 8>;
using TileSquare = DmmaTile<128, 128, 2, 4>;

struct DmmaArgs {
    Ctr128 ctr;
    PhiloxKey key;
    const double2* logtab;
    int64_t R, v0, ublk0;
    int kshift;
    int x_t;              // 1: Philox blocks run along the ROWS of X (Axis::Short operators, transposed uses): v0 counts along k
    int64_t P, Q, K;
    int steps_total, splits;
    double alpha, beta;
    const double* Y;
    int64_t ycs;          // column stride of Y in elements when rows are contiguous (yrs == 1), else the ROW stride (YMN)
    double* C;
    int64_t crs, ccs;
    const double* X;      // materialised operator (XMAT kernels): X(i, k) = X[(xr0 + i) * xld + xk0 + k]
    int64_t xld, xr0, xk0;
    int x_al;             // rows of X can be read with 16-byte loads
    double* W;            // split-K workspace W[split][j][i] (i fastest, ld = P_pad) or null
    int64_t P_pad, Q_pad;
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
    const uint32_t d = (uint32_t) __cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}

// YMN: Y is contiguous along Q (left sketch of RowMajor data, right sketch of ColMajor data); the tile is staged as
// [k][q] and the B fragments are read across rows.
// ---- the kernel: warp-specialised ------------------------------------------------------------------------------
// 12 warps: warps 0-3 (one per scheduler) only PRODUCE the S tile of every stage (generated, or copied from a materialised
// operator); warps 4-11 (two per scheduler) issue the DMMAs and, spread through the second half of each stage, the
// cp.async copies of the Y tile two steps ahead (8 per thread). A DMMA warp never leaves its DMMA stream for long, the
// producer warp of a scheduler fills the issue slots between DMMAs. Register split with setmaxnreg: producers 72, DMMA
// warps 216. History of the design and what each step measured: DESIGN.md section 4 (K2b).
#ifndef RB_DMMA_PROD_WARPS
#define RB_DMMA_PROD_WARPS 4      // producer warps per CTA: one per scheduler (8 measured: no better)
#endif
#if RB_DMMA_PROD_WARPS == 8
#define RB_DMMA_REGS_PROD "64"    // 512 threads start at 128 registers; 256 x 64 + 256 x 192 = the whole file
#define RB_DMMA_REGS_MMA "192"
#else
#define RB_DMMA_REGS_PROD "72"
#define RB_DMMA_REGS_MMA "216"
#endif
constexpr int WS_PROD = 32 * RB_DMMA_PROD_WARPS, WS_THREADS = 256 + WS_PROD;

// XMAT: the operator is materialised (S.buff != nullptr, skge.hh:174-181): the producer warps copy its DM x 16 tiles
// from global memory instead of generating them (plain loads, three stages ahead of the DMMA warps).
template <bool GAUSS, bool YMN, class TILE, bool XMAT = false>
__global__ void __launch_bounds__(WS_THREADS, 1) skge3_dmma_ws_kernel(const DmmaArgs a) {
    constexpr int DM = TILE::DM, DN = TILE::DN, D_WN = TILE::D_WN, D_MI = TILE::D_MI, D_NI = TILE::D_NI, DLQ = TILE::DLQ;
    constexpr int P_GR = DM / (WS_PROD / D_CPR);             // rows of the S tile generated per producer thread and step
    __shared__ __align__(16) double2 logtab[GAUSS ? LOGF_TABLE_ENTRIES : 1];
    extern __shared__ __align__(16) double dsm[];
    double* Xs = dsm;                              // [D_STAGES][DM][DLD]
    double* Ys = dsm + D_STAGES * DM * DLD;        // [D_STAGES][DN][DLD]
    __shared__ __align__(8) unsigned long long bars[2 * D_STAGES];
    const uint32_t bar_full = tma::smem_u32(bars), bar_empty = bar_full + 8 * D_STAGES;
    if constexpr (GAUSS) load_logf_table(logtab, a.logtab);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t i0 = (int64_t) blockIdx.y * DM, j0 = (int64_t) blockIdx.x * DN;
    const int split = blockIdx.z;
    const int per = a.steps_total / a.splits, rem = a.steps_total % a.splits;
    const int s_begin = split * per + min(split, rem);
    const int nsteps = per + (split < rem ? 1 : 0);

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < D_STAGES; ++i) {
            tma::mbar_init(bar_full + 8u * i, WS_PROD + (WS_THREADS - WS_PROD));   // S tile: producers; Y tile: cp.async of the DMMA warps
            tma::mbar_init(bar_empty + 8u * i, (WS_THREADS - WS_PROD) / 32);
        }
        tma::mbar_fence_init();
    }
    __syncthreads();                  // logtab, barriers
    auto arrive = [&](uint32_t bar) {
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
    };

    // Y tile of one step: DN x DK doubles = DN * DK / 2 16-byte copies, issued by the DMMA warps (256 threads). Thread t
    // copies chunks t, t + 256, ...: consecutive chunks of a thread are a fixed number of tile rows apart, so inside the
    // matrix (full tile, full K step -- all but the edge tiles and the last step) a copy is one pointer addition and one
    // LDGSTS; the bounds-checked form below costs ~25 instructions per copy and made the stage boundary 22% of a DMMA
    // warp's time (ncu source view, round 2).
    constexpr int Y_CPT = (DN * DK / 2) / (WS_THREADS - WS_PROD);                       // copies per thread and step
    constexpr int Y_RPQ = (WS_THREADS - WS_PROD) / (YMN ? DN / 2 : DK / 2);             // tile rows between two copies of a thread
    const int yt = (int) threadIdx.x - WS_PROD;
    const int y_r0 = yt / (YMN ? DN / 2 : DK / 2), y_c0 = (yt % (YMN ? DN / 2 : DK / 2)) * 2;
    const bool tile_full = j0 + DN <= a.Q;
    const double* y_first = YMN ? a.Y + ((int64_t) s_begin * DK + y_r0) * a.ycs + j0 + y_c0
                                : a.Y + (j0 + y_r0) * a.ycs + (int64_t) s_begin * DK + y_c0;
    const int64_t y_step = YMN ? DK * a.ycs : DK, y_q = Y_RPQ * a.ycs;
    const int y_dst0 = y_r0 * (YMN ? DLQ : DLD) + y_c0;
    auto y_is_fast = [&](int step) { return tile_full && (int64_t) (s_begin + step + 1) * DK <= a.K; };
    auto y_arrive = [&](int buf) {
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_full + 8u * (uint32_t) buf) : "memory");
    };
    // bounds-checked copies of one step (edge tiles, the last K step, the prologue when it is ragged)
    auto load_y_checked = [&](int buf, int step, int t) {
        const int64_t k0 = (int64_t) (s_begin + step) * DK;
#pragma unroll 1
        for (int q = 0; q < Y_CPT; ++q) {
            const int ch = t + (WS_THREADS - WS_PROD) * q;
            if constexpr (!YMN) {
                const int jj = ch / (DK / 2), kc = (ch % (DK / 2)) * 2;
                double* dst = Ys + ((size_t) buf * DN + jj) * DLD + kc;
                const int64_t j = j0 + jj, k = k0 + kc;
                if (j < a.Q && k + 1 < a.K) {
                    cp_async16(dst, a.Y + j * a.ycs + k, 16);
                } else {
                    double y0 = 0.0, y1 = 0.0;
                    if (j < a.Q && k < a.K) y0 = a.Y[j * a.ycs + k];
                    dst[0] = y0; dst[1] = y1;
                }
            } else {
                const int kk = ch / (DN / 2), jc = (ch % (DN / 2)) * 2;
                double* dst = Ys + (size_t) buf * DN * DLD + kk * DLQ + jc;
                const int64_t j = j0 + jc, k = k0 + kk;
                if (k < a.K && j + 1 < a.Q) {
                    cp_async16(dst, a.Y + k * a.ycs + j, 16);
                } else {
                    double y0 = 0.0, y1 = 0.0;
                    if (k < a.K && j < a.Q) y0 = a.Y[k * a.ycs + j];
                    dst[0] = y0; dst[1] = y1;
                }
            }
        }
    };
    auto load_y = [&](int buf, int step, int t) {
        if (y_is_fast(step)) {
            const double* src = y_first + (int64_t) step * y_step;
            double* dst = Ys + (size_t) buf * DN * DLD + y_dst0;
#pragma unroll
            for (int q = 0; q < Y_CPT; ++q) {
                cp_async16(dst + q * (Y_RPQ * (YMN ? DLQ : DLD)), src, 16);
                src += y_q;
            }
        } else {
            load_y_checked(buf, step, t);
        }
        y_arrive(buf);
    };

    if (warp < WS_PROD / 32) {
        // ---------------- producers ----------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " RB_DMMA_REGS_PROD ";");
        const int xc = tid % D_CPR, xr = tid / D_CPR;
        const uint64_t seed_lo = ((uint64_t) a.ctr.c1 << 32) | a.ctr.c0, seed_hi = ((uint64_t) a.ctr.c3 << 32) | a.ctr.c2;
        for (int step = 0; step < nsteps; ++step) {
            const int buf = step % D_STAGES;
            if (step >= D_STAGES) tma::mbar_wait(bar_empty + 8u * (uint32_t) buf, (uint32_t) ((step / D_STAGES - 1) & 1));
            // S tile
            if (a.x_t) {
                // blocks along the rows of X: thread -> (column k of the step, 4-row block); element (i, k) is lane
                // (u0 + i0 + i) & 3 of block (v0 + k) * R + (u0 + i0 + i) / 4, u0 % 4 == 0 (checked by the launcher)
                const int kl = tid & (DK - 1), rb0 = tid / DK;
#pragma unroll 2
                for (int rr = 0; rr < P_GR; ++rr) {
                    const int rb = rb0 + (WS_PROD / DK) * rr;
                    if constexpr (XMAT) {
                        // materialised operator contiguous along the rows of X (a filled Axis::Short operator):
                        // X(i, k) = X[(xk0 + k) * xld + xr0 + i]; the same thread mapping, four consecutive rows per load pair
                        const int64_t i = i0 + 4 * rb, k = (int64_t) (s_begin + step) * DK + kl;
                        double2 x01 = make_double2(0.0, 0.0), x23 = make_double2(0.0, 0.0);
                        if (k < a.K && i < a.P) {
                            const double* src = a.X + (a.xk0 + k) * a.xld + a.xr0 + i;
                            if (a.x_al && i + 3 < a.P) {
                                x01 = __ldg(reinterpret_cast<const double2*>(src));
                                x23 = __ldg(reinterpret_cast<const double2*>(src) + 1);
                            } else {
                                x01.x = __ldg(src);
                                if (i + 1 < a.P) x01.y = __ldg(src + 1);
                                if (i + 2 < a.P) x23.x = __ldg(src + 2);
                                if (i + 3 < a.P) x23.y = __ldg(src + 3);
                            }
                        }
                        double* dst = Xs + ((size_t) buf * DM + 4 * rb) * DLD + kl;
                        dst[0] = x01.x; dst[DLD] = x01.y; dst[2 * DLD] = x23.x; dst[3 * DLD] = x23.y;
                        continue;
                    }
                    const uint64_t o = (uint64_t) ((a.v0 + (int64_t) (s_begin + step) * DK + kl) * a.R + a.ublk0 + (i0 >> 2) + rb);
                    const uint64_t lo = seed_lo + o;
                    const uint64_t hi = seed_hi + (lo < seed_lo ? 1ull : 0ull);
                    const Ctr128 cc{(uint32_t) lo, (uint32_t) (lo >> 32), (uint32_t) hi, (uint32_t) (hi >> 32)};
                    const float4 f = transform4<GAUSS>(philox4x32_10(cc, a.key), logtab);
                    double* dst = Xs + ((size_t) buf * DM + 4 * rb) * DLD + kl;
                    dst[0] = finish_sample<double, GAUSS>(f.x);
                    dst[DLD] = finish_sample<double, GAUSS>(f.y);
                    dst[2 * DLD] = finish_sample<double, GAUSS>(f.z);
                    dst[3 * DLD] = finish_sample<double, GAUSS>(f.w);
                }
            } else
#pragma unroll 2
            for (int rr = 0; rr < P_GR; ++rr) {
                const int row = xr + (WS_PROD / D_CPR) * rr;
                if constexpr (XMAT) {
                    const int64_t i = i0 + row, k = (int64_t) (s_begin + step) * DK + 4 * xc;
                    const double* src = a.X + (a.xr0 + i) * a.xld + a.xk0 + k;
                    double2 x01 = make_double2(0.0, 0.0), x23 = make_double2(0.0, 0.0);
                    if (i < a.P) {
                        if (a.x_al && k + 3 < a.K) {
                            x01 = __ldg(reinterpret_cast<const double2*>(src));
                            x23 = __ldg(reinterpret_cast<const double2*>(src) + 1);
                        } else {
                            if (k < a.K) x01.x = __ldg(src);
                            if (k + 1 < a.K) x01.y = __ldg(src + 1);
                            if (k + 2 < a.K) x23.x = __ldg(src + 2);
                            if (k + 3 < a.K) x23.y = __ldg(src + 3);
                        }
                    }
                    double* dst = Xs + ((size_t) buf * DM + row) * DLD + 4 * xc;
                    *reinterpret_cast<double2*>(dst) = x01;
                    *reinterpret_cast<double2*>(dst + 2) = x23;
                    continue;
                }
                const uint64_t o = (uint64_t) ((a.v0 + i0 + row) * a.R + a.ublk0 + xc) + (uint64_t) D_CPR * (uint64_t) (s_begin + step);
                const uint64_t lo = seed_lo + o;
                const uint64_t hi = seed_hi + (lo < seed_lo ? 1ull : 0ull);
                const Ctr128 cc{(uint32_t) lo, (uint32_t) (lo >> 32), (uint32_t) hi, (uint32_t) (hi >> 32)};
                float4 f = transform4<GAUSS>(philox4x32_10(cc, a.key), logtab);
                if (a.kshift) {
                    const float4 h = transform4<GAUSS>(philox4x32_10(ctr_add(cc, 1), a.key), logtab);
                    if (a.kshift == 1) f = make_float4(f.y, f.z, f.w, h.x);
                    else if (a.kshift == 2) f = make_float4(f.z, f.w, h.x, h.y);
                    else f = make_float4(f.w, h.x, h.y, h.z);
                }
                double* dst = Xs + ((size_t) buf * DM + row) * DLD + 4 * xc;
                *reinterpret_cast<double2*>(dst) = make_double2(finish_sample<double, GAUSS>(f.x), finish_sample<double, GAUSS>(f.y));
                *reinterpret_cast<double2*>(dst + 2) = make_double2(finish_sample<double, GAUSS>(f.z), finish_sample<double, GAUSS>(f.w));
            }
            arrive(bar_full + 8u * (uint32_t) buf);
        }
        return;
    }
    // ---------------- DMMA warps ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 " RB_DMMA_REGS_MMA ";");
    const int cw = warp - WS_PROD / 32;
    const int wi = cw / D_WN, wj = cw % D_WN;
    const int g = lane >> 2, t4 = lane & 3;
    double acc[D_MI][D_NI][2];
#pragma unroll
    for (int mi = 0; mi < D_MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < D_NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    const int ct = tid - WS_PROD;
    // One stage = two halves of DK / 8 DMMA groups. The copies of the Y tile two steps ahead go out DURING the second half,
    // one after every few rows of DMMAs (the stage they refill was left by every DMMA warp one iteration ago, so the wait
    // in front of them does not stall): issued at the stage boundary they were time in which neither DMMA warp of a
    // scheduler fed the pipe. What remains at the boundary: one arrive, one wait, the first fragment loads.
    auto half_stage = [&](int buf, int h, auto&& after_row) {
        const double* xb = Xs + ((size_t) buf * DM + wi * (D_MI * 8) + g) * DLD + t4 + h * (DK / 2);
        const double* yb = YMN ? Ys + (size_t) buf * DN * DLD + (t4 + h * (DK / 2)) * DLQ + wj * (D_NI * 8) + g
                               : Ys + ((size_t) buf * DN + wj * (D_NI * 8) + g) * DLD + t4 + h * (DK / 2);
#pragma unroll
        for (int k4 = 0; k4 < DK / 8; ++k4) {
            double af[D_MI], bf[D_NI];
#pragma unroll
            for (int mi = 0; mi < D_MI; ++mi) af[mi] = xb[mi * 8 * DLD + k4 * 4];
#pragma unroll
            for (int ni = 0; ni < D_NI; ++ni) bf[ni] = YMN ? yb[k4 * 4 * DLQ + ni * 8] : yb[ni * 8 * DLD + k4 * 4];
#pragma unroll
            for (int mi = 0; mi < D_MI; ++mi) {
#pragma unroll
                for (int ni = 0; ni < D_NI; ++ni) dmma(acc[mi][ni], af[mi], bf[ni]);
                after_row(k4 * D_MI + mi);
            }
        }
    };
    constexpr int ROWS_HALF = (DK / 8) * D_MI;               // DMMA rows (D_NI DMMAs each) per half stage
    static_assert(ROWS_HALF % Y_CPT == 0, "copies spread evenly over the second half of a stage");
    for (int s0 = 0; s0 < D_AHEAD && s0 < nsteps; ++s0) load_y(s0, s0, ct);
    for (int step = 0; step < nsteps; ++step) {
        const int buf = step % D_STAGES;
        tma::mbar_wait(bar_full + 8u * (uint32_t) buf, (uint32_t) ((step / D_STAGES) & 1));
        half_stage(buf, 0, [](int) {});
        const int nstep = step + D_AHEAD, nb = nstep % D_STAGES;
        const bool refill = nstep < nsteps;
        if (refill && nstep >= D_STAGES) tma::mbar_wait(bar_empty + 8u * (uint32_t) nb, (uint32_t) ((nstep / D_STAGES - 1) & 1));
        if (refill && y_is_fast(nstep)) {
            const double* src = y_first + (int64_t) nstep * y_step;
            double* dst = Ys + (size_t) nb * DN * DLD + y_dst0;
            half_stage(buf, 1, [&](int row) {
                if (row % (ROWS_HALF / Y_CPT) == 0) {
                    cp_async16(dst + (row / (ROWS_HALF / Y_CPT)) * (Y_RPQ * (YMN ? DLQ : DLD)), src, 16);
                    src += y_q;
                }
            });
            y_arrive(nb);
        } else {
            if (refill) load_y(nb, nstep, ct);
            half_stage(buf, 1, [](int) {});
        }
        __syncwarp();
        if (lane == 0) arrive(bar_empty + 8u * (uint32_t) buf);
    }
#pragma unroll
    for (int mi = 0; mi < D_MI; ++mi) {
        const int64_t i = i0 + wi * (D_MI * 8) + mi * 8 + g;
#pragma unroll
        for (int ni = 0; ni < D_NI; ++ni) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t j = j0 + wj * (D_NI * 8) + ni * 8 + 2 * t4 + e;
                if (a.W) {
                    a.W[((int64_t) split * a.Q_pad + j) * a.P_pad + i] = acc[mi][ni][e];
                } else if (i < a.P && j < a.Q) {
                    double* cp = a.C + i * a.crs + j * a.ccs;
                    double r = a.alpha * acc[mi][ni][e];
                    if (a.beta != 0.0) r += a.beta * (*cp);
                    *cp = r;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) splitk_reduce_f64_kernel(const double* __restrict__ W, int splits, int64_t P, int64_t Q,
                                                                int64_t P_pad, int64_t Q_pad, double alpha, double beta,
                                                                double* __restrict__ C, int64_t crs, int64_t ccs) {
    const int64_t total = P * Q;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
        const int64_t j = e / P, i = e - j * P;
        const double* w = W + j * P_pad + i;
        double s = 0.0;
#pragma unroll 4
        for (int sp = 0; sp < splits; ++sp) s += w[(int64_t) sp * Q_pad * P_pad];
        double* cp = C + i * crs + j * ccs;
        double r = alpha * s;
        if (beta != 0.0) r += beta * (*cp);
        *cp = r;
    }
}

}  // namespace

// Gaussian operators: the producer warps' Box-Muller arithmetic needs the FP64 pipe for its DFMAs, and a DMMA stream starves
// them (tools/micro/dmma_dfma.cu: one dependent DFMA of another warp gets through every ~270 cycles while DMMAs issue back
// to back), so the fused kernel is producer-bound: 70-75 ms on the C3 shard against 65 with a Uniform operator.
// "dmma_materialise" = 1 (default) generates each K panel of op(S) once into scratch memory ("dmma_panel_mb", default 2 GB)
// with the fill kernel -- no DMMA in flight, full rate -- and runs the materialised-operator instantiation on it: 64.5 ms.
// Same operand values; panel sums are added in order (beta = 1 after the first panel), 1e-14 from the fused result.
// 0 = fused. Short problems (K < 4096) stay fused.
static int dense_dmma_f64_via_panel(const DenseProblem<double>& p, bool x_t, cudaStream_t st) {
    int64_t mb = get_option("dmma_panel_mb");
    if (mb < 16) mb = 16;
    int64_t kp = (mb << 20) / 8 / (p.P > 0 ? p.P : 1);
    kp = (kp / 1024) * 1024;
    if (kp < 1024) kp = 1024;
    if (kp > p.K) kp = (p.K + 1) / 2 * 2;
    const int64_t ld = kp;
    double* panel = (double*) workspace(5, (size_t) p.P * (size_t) ld * sizeof(double), st);
    if (!panel) return -2;                          // no room for the scratch panel: the caller runs the fused kernel
    for (int64_t k0 = 0; k0 < p.K; k0 += kp) {
        const int64_t kc = (p.K - k0 < kp) ? p.K - k0 : kp;
        int rc;
        if (!x_t) rc = launch_fill_dense<double>(p.gen, p.family, p.v0, p.P, p.u0 + k0, kc, panel, ld, 1, st);
        else rc = launch_fill_dense<double>(p.gen, p.family, p.v0 + k0, kc, p.u0, p.P, panel, 1, ld, st);
        if (rc) return rc;
        DenseProblem<double> q = p;
        q.S_buff = panel; q.S_ld = ld;
        q.v0 = 0; q.u0 = 0; q.vi = 1; q.ui = 0; q.vk = 0; q.uk = 1;
        q.K = kc;
        q.Y = p.Y + k0 * p.yrs;
        q.beta = (k0 == 0) ? p.beta : 1.0;
        rc = launch_dense_dmma_f64(q, st);
        if (rc) return rc < 0 ? fail("operator panel: the DMMA kernel refused its own panel") : rc;
    }
    return 0;
}

int launch_dense_dmma_f64(const DenseProblem<double>& p, cudaStream_t st) {
    const bool xmat = p.S_buff != nullptr;
    if (xmat && get_option("dense_path") == 4) return -1;
    if (!xmat && p.family == 'G' && !p.gen.logtab) return -1;
    // Philox blocks (rows of a materialised operator) along K, or -- generated operators only -- along the rows of X
    // (Axis::Short operators and transposed uses, dense_skops.hh:187-199) when the window starts on a block boundary
    const bool x_t = (p.ui == 1 && p.vk == 1);
    // (a materialised operator with that orientation, i.e. a filled Axis::Short one, is copied tile by tile with the same mapping)
    if (!(p.uk == 1 && p.vi == 1) && !(x_t && (xmat || (p.u0 & 3) == 0))) return -1;
    if (x_t && xmat && get_option("tc_xmn") == 0) return -1;  // experiment switch: such operators to the generic kernel
    // Y K-contiguous, or Q-contiguous (left sketch of RowMajor data, right sketch of ColMajor data)
    const bool y_mn = (p.yrs != 1);
    if (y_mn && p.ycs != 1) return -1;
    if (y_mn && get_option("dense_path") == 3) return -1;     // experiment switch: Q-contiguous data to the generic kernel
    if (p.K < 32 || p.P < 1 || p.Q < 1) return -1;
    if ((reinterpret_cast<uintptr_t>(p.Y) & 15) != 0 || ((y_mn ? p.yrs : p.ycs) & 1) != 0) return -1;   // 16-byte cp.async
    if ((int64_t) p.P * p.Q < 64 * 64 && p.K < 4096) return -1;
    const bool wide = p.Q > 128;
    const int DM = wide ? TileWide::DM : TileSquare::DM, DN = wide ? TileWide::DN : TileSquare::DN;
    const int64_t tiles_p = (p.P + DM - 1) / DM, tiles_q = (p.Q + DN - 1) / DN;
    if (tiles_p > 65535 || tiles_q > 0x7fffffff) return -1;
    const int64_t steps = (p.K + DK - 1) / DK;
    if (steps > 0x7fffffff) return -1;
    if (!xmat && p.family == 'G' && get_option("dmma_materialise") != 0 && p.K >= 4096) {
        const int prc = dense_dmma_f64_via_panel(p, x_t, st);
        if (prc != -2) return prc;                      // -2: no memory for the panel, run fused
    }
    const int64_t tiles = tiles_p * tiles_q;
    const int sms = sm_count();
    // split K only to fill the SMs: pick the split count (<= 16) with the best wave efficiency
    int splits = 1;
    {
        double best = 0;
        for (int s = 1; s <= 16; ++s) {
            if (s > 1 && steps / s < 64) break;
            const int64_t ctas = tiles * s;
            const int64_t waves = (ctas + sms - 1) / sms;
            const double eff = (double) ctas / (double) (waves * sms);
            if (eff > best + 0.03) { best = eff; splits = s; }
        }
    }
    if (get_option("tc_splits") > 0) {
        splits = (int) get_option("tc_splits");
        if (splits > steps) splits = (int) steps;
    }
    DmmaArgs a;
    a.ctr = p.gen.ctr; a.key = p.gen.key; a.R = p.gen.R; a.logtab = p.gen.logtab;
    a.v0 = p.v0;
    a.kshift = x_t ? 0 : (int) (p.u0 & 3);
    a.x_t = x_t ? 1 : 0;
    a.ublk0 = p.u0 >> 2;
    a.P = p.P; a.Q = p.Q; a.K = p.K;
    a.steps_total = (int) steps; a.splits = splits;
    a.alpha = p.alpha; a.beta = p.beta;
    a.Y = p.Y; a.ycs = y_mn ? p.yrs : p.ycs;
    a.C = p.C; a.crs = p.crs; a.ccs = p.ccs;
    a.P_pad = tiles_p * DM; a.Q_pad = tiles_q * DN;
    // K-contiguous operator: X(i, k) = X[(xr0 + i) * xld + xk0 + k]; row-contiguous (x_t): X(i, k) = X[(xk0 + k) * xld + xr0 + i]
    a.X = p.S_buff; a.xld = p.S_ld; a.xr0 = (xmat && x_t) ? p.u0 : p.v0; a.xk0 = (xmat && x_t) ? p.v0 : p.u0;
    a.x_al = (xmat && (reinterpret_cast<uintptr_t>(p.S_buff) & 15) == 0 && (p.S_ld & 1) == 0 && (p.u0 & 1) == 0) ? 1 : 0;
    a.W = nullptr;
    if (splits > 1) {
        a.W = (double*) workspace(6, (size_t) splits * a.P_pad * a.Q_pad * sizeof(double), st);
        if (!a.W) return fail_cuda(cudaErrorMemoryAllocation, "split-K workspace");
    }
    const size_t smem = wide ? TileWide::smem : TileSquare::smem;
    const bool gauss = p.family == 'G';
    dim3 grid((unsigned) tiles_q, (unsigned) tiles_p, (unsigned) splits);
    static DevOnce attr_done[20];
    auto launch_ws = [&](auto kern, DevOnce& done) -> int {
        if (done.need()) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess) {
                cudaGetLastError();
                return -1;
            }
            done.done();
        }
        kern<<<grid, WS_THREADS, smem, st>>>(a);
        return 0;
    };
    int lrc;
    if (xmat) {
        if (wide) lrc = y_mn ? launch_ws(skge3_dmma_ws_kernel<false, true, TileWide, true>, attr_done[19]) : launch_ws(skge3_dmma_ws_kernel<false, false, TileWide, true>, attr_done[18]);
        else lrc = y_mn ? launch_ws(skge3_dmma_ws_kernel<false, true, TileSquare, true>, attr_done[17]) : launch_ws(skge3_dmma_ws_kernel<false, false, TileSquare, true>, attr_done[16]);
    } else {
        if (wide) {
            if (gauss) lrc = y_mn ? launch_ws(skge3_dmma_ws_kernel<true, true, TileWide>, attr_done[15]) : launch_ws(skge3_dmma_ws_kernel<true, false, TileWide>, attr_done[14]);
            else lrc = y_mn ? launch_ws(skge3_dmma_ws_kernel<false, true, TileWide>, attr_done[13]) : launch_ws(skge3_dmma_ws_kernel<false, false, TileWide>, attr_done[12]);
        } else {
            if (gauss) lrc = y_mn ? launch_ws(skge3_dmma_ws_kernel<true, true, TileSquare>, attr_done[11]) : launch_ws(skge3_dmma_ws_kernel<true, false, TileSquare>, attr_done[10]);
            else lrc = y_mn ? launch_ws(skge3_dmma_ws_kernel<false, true, TileSquare>, attr_done[9]) : launch_ws(skge3_dmma_ws_kernel<false, false, TileSquare>, attr_done[8]);
        }
    }
    if (lrc) return -1;
    count_launch();
    count_tc_launch();
    RB_CUDA(cudaGetLastError());
    if (splits > 1) {
        int64_t gr = (p.P * p.Q + 255) / 256;
        if (gr > (int64_t) sms * 8) gr = (int64_t) sms * 8;
        splitk_reduce_f64_kernel<<<(unsigned) gr, 256, 0, st>>>(a.W, splits, p.P, p.Q, a.P_pad, a.Q_pad, p.alpha, p.beta,
                                                               p.C, p.crs, p.ccs);
        count_launch();
        RB_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace rb
