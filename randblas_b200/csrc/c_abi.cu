// extern "C" boundary of librandblas_b200.so (declared in include/randblas_b200.h).
// Argument validation mirrors the reference's randblas_require checks (cited per function); host pointers
// are staged through device memory here, device pointers are used in place.
#include <atomic>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <vector>
#include "../../include/randblas_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace rb {

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};
static std::atomic<int64_t> g_dense_path{0};
static std::atomic<int64_t> g_saso_fill_path{0};   // 1: warp-per-vector SASO fill kernel (the 64-bit-index path)
static std::atomic<int64_t> g_tc_launches{0};
static std::atomic<int64_t> g_tc_splits{0};
static std::atomic<int64_t> g_tc_halves{1};      // generator warps split into two halves, one per stage (skge3_f32_tc.cu)
static std::atomic<int64_t> g_tc_materialise{1};  // Gaussian float operators with >= 3 column tiles: generate each K panel once, then the XMAT kernel
static std::atomic<int64_t> g_saso_bin_path{0};    // binning pass of the SASO apply: 0 thread per column where possible, 1 lane per entry
static std::atomic<int64_t> g_saso_rows{1};        // SASO apply: 1 (default) a lane owns a whole row of the tile, 0 an 8-lane group owns 8 rows
static std::atomic<int64_t> g_fill_unroll{1};      // Uniform float fill of long vectors: 1 = 16 Philox blocks per thread and tile, 0 = 4
static std::atomic<int64_t> g_fill_rep{1};         // Gaussian fill of long vectors: 1 = tiles of 4 passes of 8 blocks per thread, 0 = one pass
static std::atomic<int64_t> g_tc_xmn{1};             // row-contiguous MATERIALISED operators (filled Axis::Short): 1 tensor cores / DMMA, 0 generic kernel
static std::atomic<int64_t> g_tc_ymn{0};             // float tensor-core kernel, Q-contiguous data: 0 MN-major operand, 1 transposing path
static std::atomic<int64_t> g_dmma_materialise{1};  // double Gaussian operators: panel-materialise + XMAT DMMA kernel
static std::atomic<int64_t> g_dmma_panel_mb{2048};   // ... size of that panel
static std::atomic<int64_t> g_tc_pair{1};        // CTA pairs (cta_group::2) in the float tensor-core kernel where the shape allows
static std::atomic<int64_t> g_tc_cluster{1};     // 2-CTA clusters sharing the generated operator tile (skge3_f32_tc.cu)
static std::atomic<int64_t> g_spdata_path{0};   // 0 auto (k-group kernel), 1 force the column-owner kernel (no atomics)
static std::atomic<int64_t> g_h2d_chunk_mb{64};   // block size of the host-pointer sketch pipeline (MB of A per block)
static std::atomic<int64_t> g_saso_path{0};     // 0 auto, 1 force the atomic kernel, 2 force the binned kernel

void set_error(const std::string& m) { g_err = m; }
int fail(const std::string& m) { g_err = m; return RB_ERR_ARG; }
int fail_cuda(cudaError_t e, const char* what) {
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return RB_ERR_CUDA;
}
void count_launch(int n) { g_launches += n; }
void count_tc_launch() { g_tc_launches += 1; }
static std::atomic<int64_t> g_owner_launches{0};
void count_owner_launch() { g_owner_launches += 1; }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

const double2* logf_table_device() {
    static std::mutex mu;
    static double2* tabs[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!tabs[dev]) {
        std::vector<double2> h(LOGF_TABLE_ENTRIES);
        for (int k = -33; k <= 0; ++k)
            for (int i = 0; i < 16; ++i) {
                h[(k + 33) * 16 + i].x = std::ldexp(h_logf_tab[2 * i], 896);      // invc * 2^896 (philox.cuh: logf_exact)
                h[(k + 33) * 16 + i].y = std::fma((double) k, 0x1.62e42fefa39efp-1, h_logf_tab[2 * i + 1]);
            }
        const double hpi = 0x1.921fb54442d18p+0;
        double* q = reinterpret_cast<double*>(h.data()) + QUADRANT_TABLE_OFFSET;     // -n pi/2, n = -2..2
        q[0] = 2 * hpi; q[1] = hpi; q[2] = 0.0; q[3] = -hpi; q[4] = -2 * hpi; q[5] = 0.0;
        double2* d = nullptr;
        if (cudaMalloc(&d, h.size() * sizeof(double2)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (cudaMemcpy(d, h.data(), h.size() * sizeof(double2), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaGetLastError(); cudaFree(d); return nullptr;
        }
        tabs[dev] = d;
    }
    return tabs[dev];
}

int64_t get_option(const char* name) {
    if (!std::strcmp(name, "dense_path")) return g_dense_path.load();
    if (!std::strcmp(name, "saso_fill_path")) return g_saso_fill_path.load();
    if (!std::strcmp(name, "tc_splits")) return g_tc_splits.load();
    if (!std::strcmp(name, "tc_cluster")) return g_tc_cluster.load();
    if (!std::strcmp(name, "tc_pair")) return g_tc_pair.load();
    if (!std::strcmp(name, "tc_ymn")) return g_tc_ymn.load();
    if (!std::strcmp(name, "tc_xmn")) return g_tc_xmn.load();
    if (!std::strcmp(name, "fill_rep")) return g_fill_rep.load();
    if (!std::strcmp(name, "fill_unroll")) return g_fill_unroll.load();
    if (!std::strcmp(name, "saso_rows")) return g_saso_rows.load();
    if (!std::strcmp(name, "saso_bin_path")) return g_saso_bin_path.load();
    if (!std::strcmp(name, "dmma_materialise")) return g_dmma_materialise.load();
    if (!std::strcmp(name, "dmma_panel_mb")) return g_dmma_panel_mb.load();
    if (!std::strcmp(name, "tc_materialise")) return g_tc_materialise.load();
    if (!std::strcmp(name, "tc_halves")) return g_tc_halves.load();
    if (!std::strcmp(name, "saso_path")) return g_saso_path.load();
    if (!std::strcmp(name, "h2d_chunk_mb")) return g_h2d_chunk_mb.load();
    if (!std::strcmp(name, "spdata_path")) return g_spdata_path.load();
    return 0;
}

// ---- cached workspace: grow-only buffers owned by one (device, stream, host thread) ----
// Every call's intermediates (split-K partials, bucket arrays, scan scratch, staging) live in buffers keyed by the
// device, the stream the call is ordered on and the calling host thread, so calls on different streams or from
// different threads never share scratch memory (the reference is re-entrant; SURVEY.md section 8b "Threading").
// Calls on ONE stream from ONE thread reuse the same buffers, which is safe because they are stream-ordered.
struct WsKey {
    int dev; cudaStream_t st; std::thread::id tid;
    bool operator<(const WsKey& o) const {
        if (dev != o.dev) return dev < o.dev;
        if (st != o.st) return st < o.st;
        return tid < o.tid;
    }
};
struct WsSlot { void* p = nullptr; size_t bytes = 0; };
struct WsSet { WsSlot s[RB_WS_SLOTS]; };
static std::mutex g_ws_mu;
static std::map<WsKey, WsSet> g_ws;

void* workspace(int slot, size_t bytes, cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || slot < 0 || slot >= RB_WS_SLOTS) { cudaGetLastError(); return nullptr; }
    if (bytes == 0) bytes = 16;
    WsSlot* s;
    {
        std::lock_guard<std::mutex> lk(g_ws_mu);
        s = &g_ws[WsKey{dev, st, std::this_thread::get_id()}].s[slot];     // std::map: references stay valid
    }
    if (s->bytes < bytes) {
        // only this thread's earlier calls on this stream can still be using the old buffer
        if (s->p) { cudaStreamSynchronize(st); cudaFree(s->p); s->p = nullptr; s->bytes = 0; }
        size_t want = bytes + bytes / 8;
        if (cudaMalloc(&s->p, want) != cudaSuccess) {
            cudaGetLastError();
            if (cudaMalloc(&s->p, bytes) != cudaSuccess) { cudaGetLastError(); s->p = nullptr; return nullptr; }
            want = bytes;
        }
        s->bytes = want;
    }
    return s->p;
}
// Frees every cached buffer of every device (no call of the library may be in flight on another thread).
void release_workspace() {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : g_ws) {
        bool any = false;
        for (auto& sl : kv.second.s) any = any || sl.p;
        if (!any) continue;
        cudaSetDevice(kv.first.dev);
        cudaDeviceSynchronize();
        for (auto& sl : kv.second.s) if (sl.p) { cudaFree(sl.p); sl = WsSlot(); }
    }
    g_ws.clear();
    cudaSetDevice(cur);
    cudaGetLastError();
}

// Copy stream + events of the host-destination fill_dense pipeline, cached per (device, host thread).
struct CopyPipe { cudaStream_t st = nullptr; cudaEvent_t gen_done[2] = {nullptr, nullptr}, copy_done[2] = {nullptr, nullptr}; };
static CopyPipe* copy_pipe() {
    static std::mutex mu;
    static std::map<std::pair<int, std::thread::id>, CopyPipe> pipes;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> lk(mu);
    CopyPipe& c = pipes[{dev, std::this_thread::get_id()}];
    if (!c.st) {
        if (cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); c.st = nullptr; return nullptr; }
        for (int i = 0; i < 2; ++i) {
            cudaEventCreateWithFlags(&c.gen_done[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&c.copy_done[i], cudaEventDisableTiming);
        }
    }
    return &c;
}

// ---- host/device pointer handling ----
static bool on_device(const void* p) {
    if (!p) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Staging buffers come from the device's default memory pool. Its default release threshold is 0, i.e. every
// synchronisation hands the memory back to the driver and the next host-pointer call pays for a fresh allocation of
// hundreds of megabytes; keep freed blocks in the pool instead (once per device).
static void keep_pool_memory() {
    static std::atomic<uint64_t> done{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return; }
    if (done.load() & (1ull << dev)) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
    done.fetch_or(1ull << dev);
}

// A (outer x inner) strided matrix or a flat array that may live on the host. Staged copies keep `ld`.
struct Staged {
    void* dev = nullptr;
    void* host = nullptr;
    size_t elem = 0, outer = 0, inner = 0, ld = 0;
    bool staged = false, copy_back = false;
    cudaStream_t st = nullptr;

    int open(const void* p, size_t elem_, int64_t outer_, int64_t inner_, int64_t ld_, bool copy_in, bool copy_back_,
             cudaStream_t st_) {
        elem = elem_; outer = (size_t) outer_; inner = (size_t) inner_; ld = (size_t) ld_; st = st_;
        copy_back = copy_back_;
        if (!p || outer_ <= 0 || inner_ <= 0) { dev = const_cast<void*>(p); return 0; }
        if (on_device(p)) { dev = const_cast<void*>(p); return 0; }
        host = const_cast<void*>(p);
        staged = true;
        const size_t bytes = ((outer - 1) * ld + inner) * elem;
        keep_pool_memory();
        RB_CUDA(cudaMallocAsync(&dev, bytes, st));
        if (copy_in)
            RB_CUDA(cudaMemcpy2DAsync(dev, ld * elem, host, ld * elem, inner * elem, outer, cudaMemcpyHostToDevice, st));
        return 0;
    }
    int close() {
        if (!staged) return 0;
        staged = false;
        if (copy_back)
            RB_CUDA(cudaMemcpy2DAsync(host, ld * elem, dev, ld * elem, inner * elem, outer, cudaMemcpyDeviceToHost, st));
        RB_CUDA(cudaFreeAsync(dev, st));
        RB_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    ~Staged() {
        if (staged) { cudaFreeAsync(dev, st); cudaStreamSynchronize(st); }
    }
};

static inline bool ok_layout(char l) { return l == 'R' || l == 'C'; }
static inline bool ok_op(char o) { return o == 'N' || o == 'T'; }

// ---------------------------------------------------------------------------------------------
// fill_dense
template <typename T>
static int fill_dense_impl(char layout, int64_t D_rows, int64_t D_cols, char family, char axis, int64_t n_rows,
                           int64_t n_cols, int64_t ro_s, int64_t co_s, T* buff, int64_t ld, const uint32_t* ctr,
                           const uint32_t* key, uint32_t* next_ctr, void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(ok_layout(layout));
    DenseDistInfo D = make_dense_dist(D_rows, D_cols, family, axis);
    RB_REQUIRE(D.ok);                              // DenseDist ctor: n_rows > 0, n_cols > 0 (dense_skops.hh:327-328)
    RB_REQUIRE(ctr != nullptr && key != nullptr);
    RB_REQUIRE(n_rows >= 0 && n_cols >= 0 && ro_s >= 0 && co_s >= 0);
    RB_REQUIRE(D.n_rows >= n_rows + ro_s);         // dense_skops.hh:566
    RB_REQUIRE(D.n_cols >= n_cols + co_s);         // dense_skops.hh:567
    const int64_t inner = (layout == 'R') ? n_cols : n_rows, outer = (layout == 'R') ? n_rows : n_cols;
    if (ld <= 0) ld = inner;
    RB_REQUIRE(ld >= inner);                       // dense_skops.hh:101
    const DenseGen g = make_dense_gen(D, ctr, key);
    const bool nat_row = g.nat_row;
    const int64_t v0 = nat_row ? ro_s : co_s, nv = nat_row ? n_rows : n_cols;
    const int64_t u0 = nat_row ? co_s : ro_s, nu = nat_row ? n_cols : n_rows;
    // state returned by fill_dense_submat_impl (dense_skops.hh:127-130, 167-169)
    store_ctr(ctr_add(g.ctr, (uint64_t) (v0 * g.R + u0 / 4 + nv * g.R)), next_ctr);
    if (n_rows == 0 || n_cols == 0) return 0;
    RB_REQUIRE(buff != nullptr);
    const int64_t irs = (layout == 'R') ? ld : 1, ics = (layout == 'R') ? 1 : ld;
    const int64_t sv = nat_row ? irs : ics, su = nat_row ? ics : irs;
    if (on_device(buff)) return launch_fill_dense<T>(g, family, v0, nv, u0, nu, buff, sv, su, st);

    // Host destination: generate chunks of `outer` slices into two device staging buffers and copy them
    // back while the next chunk is generated.
    const size_t slice_bytes = (size_t) ld * sizeof(T);
    int64_t chunk = (int64_t) ((256ull << 20) / slice_bytes);
    if (chunk < 1) chunk = 1;
    if (chunk > outer) chunk = outer;
    T* stage[2];
    stage[0] = (T*) workspace(0, (size_t) chunk * slice_bytes, st);
    stage[1] = (outer > chunk) ? (T*) workspace(1, (size_t) chunk * slice_bytes, st) : stage[0];
    if (!stage[0] || !stage[1]) return fail_cuda(cudaErrorMemoryAllocation, "fill_dense staging");
    CopyPipe* pipe = copy_pipe();
    if (!pipe) return fail_cuda(cudaErrorUnknown, "fill_dense copy stream");
    cudaStream_t copy_st = pipe->st;
    cudaEvent_t* gen_done = pipe->gen_done;
    cudaEvent_t* copy_done = pipe->copy_done;
    int rc = 0, it = 0;
    const bool outer_is_v = (layout == 'R') == nat_row;   // does a slice of the destination hold one vector?
    for (int64_t o = 0; o < outer && rc == 0; o += chunk, ++it) {
        const int b = it & 1;
        const int64_t oc = (outer - o < chunk) ? outer - o : chunk;
        if (it >= 2) cudaStreamWaitEvent(st, copy_done[b], 0);
        if (outer_is_v) rc = launch_fill_dense<T>(g, family, v0 + o, oc, u0, nu, stage[b], sv, su, st);
        else rc = launch_fill_dense<T>(g, family, v0, nv, u0 + o, oc, stage[b], sv, su, st);
        if (rc) break;
        cudaEventRecord(gen_done[b], st);
        cudaStreamWaitEvent(copy_st, gen_done[b], 0);
        cudaError_t e = cudaMemcpy2DAsync(buff + o * ld, slice_bytes, stage[b], slice_bytes, (size_t) inner * sizeof(T),
                                          (size_t) oc, cudaMemcpyDeviceToHost, copy_st);
        if (e != cudaSuccess) { rc = fail_cuda(e, "fill_dense D2H"); break; }
        cudaEventRecord(copy_done[b], copy_st);
    }
    cudaStreamSynchronize(copy_st);
    cudaStreamSynchronize(st);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// dense operator, canonical problem builder
struct OpWindow { int64_t v0, u0; int vi, ui, vk, uk; };
// X = op(S[ro:, co:]) with element (i,k); nat_row = natural layout RowMajor
static OpWindow op_window(bool nat_row, char opS, int64_t ro, int64_t co) {
    OpWindow w;
    w.v0 = nat_row ? ro : co;
    w.u0 = nat_row ? co : ro;
    const bool u_along_k = (opS == 'N') == nat_row;
    w.vi = u_along_k ? 1 : 0; w.uk = u_along_k ? 1 : 0;
    w.ui = u_along_k ? 0 : 1; w.vk = u_along_k ? 0 : 1;
    return w;
}

template <typename T>
static int run_dense(DenseProblem<T>& p, cudaStream_t st) {
    if (g_dense_path.load() == 0 && p.P > 0 && p.Q > 0 && p.K > 0 && p.alpha != (T) 0) {
        int rc;
        if constexpr (sizeof(T) == 4) rc = launch_dense_tc_f32(p, st);
        else rc = launch_dense_dmma_f64(p, st);
        if (rc >= 0) return rc;
    }
    return launch_dense_generic<T>(p, st);
}

// side_left: B(d x n) = alpha op(S)(d x m) op(A)(m x n) + beta B          [skge.hh:154-202]
// right    : B(m x d) = alpha op(A)(m x n) op(S)(n x d) + beta B          [skge.hh:307-355]
template <typename T>
static int skge3_impl(bool left, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, T alpha,
                      int64_t D_rows, int64_t D_cols, char family, char axis, const uint32_t* ctr, const uint32_t* key,
                      const T* S_buff, int64_t ro_s, int64_t co_s, const T* A, int64_t lda, T beta, T* B, int64_t ldb,
                      void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(ok_layout(layout) && ok_op(opS) && ok_op(opA));
    RB_REQUIRE(d >= 0 && n >= 0 && m >= 0 && ro_s >= 0 && co_s >= 0);
    DenseDistInfo D = make_dense_dist(D_rows, D_cols, family, axis);
    RB_REQUIRE(D.ok);
    RB_REQUIRE(ctr != nullptr && key != nullptr);
    // dims of the operator window before op(): left (d, m), right (n, d)   [skge.hh:171, 324]
    const int64_t a1 = left ? d : n, a2 = left ? m : d;
    const int64_t rows_sub = (opS == 'N') ? a1 : a2, cols_sub = (opS == 'N') ? a2 : a1;
    RB_REQUIRE(D.n_rows >= rows_sub + ro_s);       // skge.hh:183 / dense_skops.hh:678
    RB_REQUIRE(D.n_cols >= cols_sub + co_s);       // skge.hh:184 / dense_skops.hh:679
    const int64_t rows_A = (opA == 'N') ? m : n, cols_A = (opA == 'N') ? n : m;   // op(A) is m x n on both sides
    const int64_t rows_B = left ? d : m, cols_B = left ? n : d;
    if (layout == 'C') { RB_REQUIRE(lda >= rows_A); RB_REQUIRE(ldb >= rows_B); }   // skge.hh:186-188, 339-341
    else               { RB_REQUIRE(lda >= cols_A); RB_REQUIRE(ldb >= cols_B); }   // skge.hh:189-192, 342-345
    if (rows_B == 0 || cols_B == 0) return 0;
    RB_REQUIRE(B != nullptr);
    RB_REQUIRE(A != nullptr || m == 0 || n == 0);

    Staged sA, sB, sS;
    const int64_t A_outer = (layout == 'C') ? cols_A : rows_A, A_inner = (layout == 'C') ? rows_A : cols_A;
    const int64_t B_outer = (layout == 'C') ? cols_B : rows_B, B_inner = (layout == 'C') ? rows_B : cols_B;
    const int64_t Ktot = left ? m : n;                       // contraction length
    // Host-resident A with a generated operator: stream A through two staging buffers in blocks along the
    // contraction dimension (the reference's own blocked form, skge.hh:174-181: block k0 uses the operator window
    // shifted by k0 and accumulates with beta = 1), so the host->device copy of block i+1 overlaps the kernel of block i.
    const bool chunked = A != nullptr && S_buff == nullptr && !on_device(A) && Ktot > 0 &&
                         (size_t) rows_A * (size_t) cols_A * sizeof(T) >= ((size_t) g_h2d_chunk_mb.load() << 20) / 2;
    int rc = 0;
    if (!chunked) { rc = sA.open(A, sizeof(T), A_outer, A_inner, lda, true, false, st); if (rc) return rc; }
    rc = sB.open(B, sizeof(T), B_outer, B_inner, ldb, beta != (T) 0, true, st);
    if (rc) return rc;
    rc = sS.open(S_buff, sizeof(T), D.dim_minor, D.dim_major, D.dim_major, true, false, st);
    if (rc) return rc;

    const DenseGen gen = make_dense_gen(D, ctr, key);
    const int64_t brs = (layout == 'C') ? 1 : ldb, bcs = (layout == 'C') ? ldb : 1;
    // canonical problem for the contraction block [k0, k0 + kc) with A (or its staged block) at Adev / lda_dev
    auto problem = [&](const T* Adev, int64_t lda_dev, int64_t k0, int64_t kc, T beta_blk) {
        DenseProblem<T> p;
        p.alpha = alpha; p.beta = beta_blk;
        p.gen = gen;
        p.family = family;
        p.S_buff = (const T*) sS.dev;
        p.S_ld = D.dim_major;
        const int64_t ars = (layout == 'C') ? 1 : lda_dev, acs = (layout == 'C') ? lda_dev : 1;
        p.Y = Adev;
        p.C = (T*) sB.dev;
        OpWindow w;
        if (left) {
            p.P = d; p.Q = n; p.K = kc;
            w = op_window(p.gen.nat_row, opS, ro_s, co_s);
            p.yrs = (opA == 'N') ? ars : acs; p.ycs = (opA == 'N') ? acs : ars;
            p.crs = brs; p.ccs = bcs;
        } else {
            // transpose the problem: B^T (d x m) = op(S)^T (d x n) * op(A)^T (n x m)
            p.P = d; p.Q = m; p.K = kc;
            w = op_window(p.gen.nat_row, opS == 'N' ? 'T' : 'N', ro_s, co_s);
            p.yrs = (opA == 'N') ? acs : ars; p.ycs = (opA == 'N') ? ars : acs;
            p.crs = bcs; p.ccs = brs;
        }
        p.v0 = w.v0 + k0 * w.vk; p.u0 = w.u0 + k0 * w.uk; p.vi = w.vi; p.ui = w.ui; p.vk = w.vk; p.uk = w.uk;
        return p;
    };

    if (!chunked) {
        DenseProblem<T> p = problem((const T*) sA.dev, lda, 0, Ktot, beta);
        rc = run_dense<T>(p, st);
    } else {
        const bool k_on_rows = left ? (opA == 'N') : (opA == 'T');          // contraction index runs over rows of A?
        const bool k_inner = (layout == 'C') == k_on_rows;                  // ... and is it the contiguous index?
        const int64_t other = k_inner ? A_outer : A_inner;                  // extent of the non-contracted index
        const size_t target = (size_t) g_h2d_chunk_mb.load() << 20;
        int64_t kc_max = (int64_t) (target / ((size_t) other * sizeof(T)));
        kc_max = (kc_max / 1024) * 1024;
        if (kc_max < 1024) kc_max = 1024;
        if (kc_max > Ktot) kc_max = Ktot;
        const int64_t pad = 16 / (int64_t) sizeof(T);
        const int64_t ld_stage = k_inner ? ((kc_max + pad - 1) / pad) * pad : ((A_inner + pad - 1) / pad) * pad;
        const size_t stage_bytes = (size_t) ld_stage * (size_t) (k_inner ? A_outer : kc_max) * sizeof(T);
        T* stage[2] = {(T*) workspace(8, stage_bytes, st), (T*) workspace(9, stage_bytes, st)};
        CopyPipe* pipe = copy_pipe();
        if (!stage[0] || !stage[1] || !pipe) rc = fail_cuda(cudaErrorMemoryAllocation, "sketch staging");
        // the first two blocks are a quarter and a half of the block size: the first kernel starts after a short copy
        auto block_len = [&](int i) {
            int64_t k = (i == 0) ? kc_max / 4 : (i == 1) ? kc_max / 2 : kc_max;
            k = (k / 1024) * 1024;
            return k < 1024 ? (kc_max < 1024 ? kc_max : (int64_t) 1024) : k;
        };
        int it = 0;
        int64_t kc = 0;
        for (int64_t k0 = 0; k0 < Ktot && rc == 0; k0 += kc, ++it) {
            const int b = it & 1;
            kc = block_len(it);
            if (Ktot - k0 < kc) kc = Ktot - k0;
            if (it >= 2) cudaStreamWaitEvent(pipe->st, pipe->gen_done[b], 0);      // kernel it-2 has released the buffer
            cudaError_t e;
            if (k_inner) e = cudaMemcpy2DAsync(stage[b], (size_t) ld_stage * sizeof(T), A + k0, (size_t) lda * sizeof(T),
                                               (size_t) kc * sizeof(T), (size_t) A_outer, cudaMemcpyHostToDevice, pipe->st);
            else e = cudaMemcpy2DAsync(stage[b], (size_t) ld_stage * sizeof(T), A + k0 * lda, (size_t) lda * sizeof(T),
                                       (size_t) A_inner * sizeof(T), (size_t) kc, cudaMemcpyHostToDevice, pipe->st);
            if (e != cudaSuccess) { rc = fail_cuda(e, "sketch H2D"); break; }
            cudaEventRecord(pipe->copy_done[b], pipe->st);
            cudaStreamWaitEvent(st, pipe->copy_done[b], 0);
            DenseProblem<T> p = problem(stage[b], ld_stage, k0, kc, it == 0 ? beta : (T) 1);
            rc = run_dense<T>(p, st);
            cudaEventRecord(pipe->gen_done[b], st);
        }
        if (pipe) cudaStreamSynchronize(pipe->st);
    }
    int rc2 = sS.close(); if (!rc) rc = rc2;
    rc2 = sA.close(); if (!rc) rc = rc2;
    rc2 = sB.close(); if (!rc) rc = rc2;
    return rc;
}

// ---------------------------------------------------------------------------------------------
// SASO generation
static int fill_sparse_impl(int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t* ctr, const uint32_t* key,
                            void* vals, int val_bytes, void* rows, void* cols, int idx_bytes, int64_t* nnz,
                            uint32_t* next_ctr, void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    SparseDistInfo D = make_sparse_dist(D_rows, D_cols, vec_nnz, 'S');
    RB_REQUIRE(D.ok);                          // sparse_skops.hh:219-223
    RB_REQUIRE(ctr != nullptr && key != nullptr);
    RB_REQUIRE(idx_bytes == 4 || idx_bytes == 8);
    RB_REQUIRE(val_bytes == 4 || val_bytes == 8);
    RB_REQUIRE(rows != nullptr);               // sparse_skops.hh:597
    RB_REQUIRE(cols != nullptr);               // sparse_skops.hh:598
    RB_REQUIRE(vals != nullptr);               // sparse_skops.hh:599
    if (idx_bytes == 4) RB_REQUIRE(D.dim_minor <= 2147483647LL && D.dim_major <= 2147483647LL);
    const Ctr128 c = load_ctr(ctr);
    store_ctr(ctr_add(c, (uint64_t) D.full_nnz), next_ctr);      // sparse_skops.hh:104-105
    if (nnz) *nnz = D.full_nnz;                                  // sparse_skops.hh:532
    Staged sv, sr, sc;
    int rc = sv.open(vals, (size_t) val_bytes, 1, D.full_nnz, D.full_nnz, false, true, st); if (rc) return rc;
    rc = sr.open(rows, (size_t) idx_bytes, 1, D.full_nnz, D.full_nnz, false, true, st); if (rc) return rc;
    rc = sc.open(cols, (size_t) idx_bytes, 1, D.full_nnz, D.full_nnz, false, true, st); if (rc) return rc;
    // short-axis index array is `rows` when n_rows <= n_cols (sparse_skops.hh:524-525)
    void* idx_short = (D_rows <= D_cols) ? sr.dev : sc.dev;
    void* idx_long = (D_rows <= D_cols) ? sc.dev : sr.dev;
    rc = launch_saso(c, PhiloxKey{key[0], key[1]}, vec_nnz, D.dim_major, D.dim_minor, idx_short, idx_long, idx_bytes,
                     sv.dev, val_bytes, st);
    int rc2 = sv.close(); if (!rc) rc = rc2;
    rc2 = sr.close(); if (!rc) rc = rc2;
    rc2 = sc.close(); if (!rc) rc = rc2;
    return rc;
}

// LASO generation (Axis::Long): fill_sparse_unpacked_nosub, sparse_skops.hh:534-564
static int fill_sparse_laso_impl(int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t* ctr, const uint32_t* key,
                                 void* vals, int val_bytes, void* rows, void* cols, int idx_bytes, int64_t* nnz,
                                 uint32_t* next_ctr, void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    SparseDistInfo D = make_sparse_dist(D_rows, D_cols, vec_nnz, 'L');
    RB_REQUIRE(D.ok);                          // sparse_skops.hh:219-223
    RB_REQUIRE(ctr != nullptr && key != nullptr);
    RB_REQUIRE(idx_bytes == 4 || idx_bytes == 8);
    RB_REQUIRE(val_bytes == 4 || val_bytes == 8);
    RB_REQUIRE(rows != nullptr);               // sparse_skops.hh:597
    RB_REQUIRE(cols != nullptr);               // sparse_skops.hh:598
    RB_REQUIRE(vals != nullptr);               // sparse_skops.hh:599
    RB_REQUIRE(nnz != nullptr);
    if (idx_bytes == 4) RB_REQUIRE(D.dim_minor <= 2147483647LL && D.dim_major <= 2147483647LL);
    const Ctr128 c = load_ctr(ctr);
    store_ctr(ctr_add(c, (uint64_t) (D.dim_minor * ((vec_nnz + 1) / 2))), next_ctr);      // sparse_skops.hh:274-279
    Staged sv, sr, sc;
    int rc = sv.open(vals, (size_t) val_bytes, 1, D.full_nnz, D.full_nnz, false, true, st); if (rc) return rc;
    rc = sr.open(rows, (size_t) idx_bytes, 1, D.full_nnz, D.full_nnz, false, true, st); if (rc) return rc;
    rc = sc.open(cols, (size_t) idx_bytes, 1, D.full_nnz, D.full_nnz, false, true, st); if (rc) return rc;
    // short-axis index array is `rows` when n_rows <= n_cols (sparse_skops.hh:524-525)
    void* idx_short = (D_rows <= D_cols) ? sr.dev : sc.dev;
    void* idx_long = (D_rows <= D_cols) ? sc.dev : sr.dev;
    rc = launch_laso(c, PhiloxKey{key[0], key[1]}, vec_nnz, D.dim_major, D.dim_minor, idx_long, idx_short, idx_bytes, sv.dev,
                     val_bytes, nnz, st);
    int rc2 = sv.close(); if (!rc) rc = rc2;
    rc2 = sr.close(); if (!rc) rc = rc2;
    rc2 = sc.close(); if (!rc) rc = rc2;
    return rc;
}

// sparse operator applied to dense data [skge.hh:465-492, 598-626; spmm_dispatch.hh:52-219]
template <typename T>
static int skges_impl(bool left, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, T alpha,
                      int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t* ctr, const uint32_t* key,
                      int64_t ro_s, int64_t co_s, const T* A, int64_t lda, T beta, T* B, int64_t ldb, void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(ok_layout(layout) && ok_op(opS) && ok_op(opA));
    RB_REQUIRE(d >= 0 && n >= 0 && m >= 0 && ro_s >= 0 && co_s >= 0);
    SparseDistInfo D = make_sparse_dist(D_rows, D_cols, vec_nnz, 'S');
    RB_REQUIRE(D.ok);
    RB_REQUIRE(ctr != nullptr && key != nullptr);
    const int64_t a1 = left ? d : n, a2 = left ? m : d;
    const int64_t rs = (opS == 'N') ? a1 : a2, cs = (opS == 'N') ? a2 : a1;
    RB_REQUIRE(D.n_rows >= rs + ro_s);             // coo view window, spmm_dispatch.hh:96-97
    RB_REQUIRE(D.n_cols >= cs + co_s);
    const int64_t rows_A = (opA == 'N') ? m : n, cols_A = (opA == 'N') ? n : m;
    const int64_t rows_B = left ? d : m, cols_B = left ? n : d;
    if (layout == 'C') { RB_REQUIRE(lda >= rows_A); RB_REQUIRE(ldb >= rows_B); }   // spmm_dispatch.hh:126-128
    else               { RB_REQUIRE(lda >= cols_A); RB_REQUIRE(ldb >= cols_B); }   // spmm_dispatch.hh:131-133
    if (rows_B == 0 || cols_B == 0) return 0;
    RB_REQUIRE(B != nullptr);
    RB_REQUIRE(A != nullptr || m == 0 || n == 0);

    Staged sA, sB;
    const int64_t A_outer = (layout == 'C') ? cols_A : rows_A, A_inner = (layout == 'C') ? rows_A : cols_A;
    const int64_t B_outer = (layout == 'C') ? cols_B : rows_B, B_inner = (layout == 'C') ? rows_B : cols_B;
    int rc = sA.open(A, sizeof(T), A_outer, A_inner, lda, true, false, st); if (rc) return rc;
    rc = sB.open(B, sizeof(T), B_outer, B_inner, ldb, beta != (T) 0, true, st); if (rc) return rc;

    const int64_t ars = (layout == 'C') ? 1 : lda, acs = (layout == 'C') ? lda : 1;
    const int64_t brs = (layout == 'C') ? 1 : ldb, bcs = (layout == 'C') ? ldb : 1;
    SasoProblem<T> p;
    p.alpha = alpha; p.beta = beta;
    p.ctr = load_ctr(ctr); p.key = PhiloxKey{key[0], key[1]};
    p.vec_nnz = vec_nnz; p.dim_major = D.dim_major; p.dim_minor = D.dim_minor;
    p.major_is_rows = (D_rows <= D_cols) ? 1 : 0;
    p.ro_s = ro_s; p.co_s = co_s; p.rs = rs; p.cs = cs;
    p.Y = (const T*) sA.dev; p.C = (T*) sB.dev;
    if (left) {
        p.P = d; p.Q = n; p.K = m;
        p.x_is_transposed = (opS == 'T');
        p.yrs = (opA == 'N') ? ars : acs; p.ycs = (opA == 'N') ? acs : ars;
        p.crs = brs; p.ccs = bcs;
    } else {
        p.P = d; p.Q = m; p.K = n;
        p.x_is_transposed = (opS == 'N');
        p.yrs = (opA == 'N') ? acs : ars; p.ycs = (opA == 'N') ? ars : acs;
        p.crs = bcs; p.ccs = brs;
    }
    rc = launch_saso_apply<T>(p, st);
    if (rc == -1) {
        // vec_nnz > 32: sample the operator into workspace COO arrays (int64) and use the COO kernel
        int64_t* maj = (int64_t*) workspace(3, (size_t) D.full_nnz * 8, st);
        int64_t* mnr = (int64_t*) workspace(4, (size_t) D.full_nnz * 8, st);
        T* vv = (T*) workspace(5, (size_t) D.full_nnz * sizeof(T), st);
        if (!maj || !mnr || !vv) rc = fail_cuda(cudaErrorMemoryAllocation, "SASO workspace");
        else rc = launch_saso(p.ctr, p.key, vec_nnz, D.dim_major, D.dim_minor, maj, mnr, 8, vv, (int) sizeof(T), st);
        if (!rc) {
            CooProblem<T> c;
            c.P = p.P; c.Q = p.Q; c.K = p.K; c.alpha = alpha; c.beta = beta; c.nnz = D.full_nnz; c.vals = vv;
            c.rows = p.major_is_rows ? maj : mnr; c.cols = p.major_is_rows ? mnr : maj; c.idx_bytes = 8;
            c.x_is_transposed = p.x_is_transposed; c.ro_s = ro_s; c.co_s = co_s; c.rs = rs; c.cs = cs;
            c.Y = p.Y; c.yrs = p.yrs; c.ycs = p.ycs; c.C = p.C; c.crs = p.crs; c.ccs = p.ccs;
            rc = launch_coo_apply<T>(c, st);
        }
    }
    int rc2 = sA.close(); if (!rc) rc = rc2;
    rc2 = sB.close(); if (!rc) rc = rc2;
    return rc;
}

template <typename T>
static int coo_apply_impl(int left, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, T alpha,
                          int64_t S_rows, int64_t S_cols, int64_t nnz, const T* vals, const void* rows, const void* cols,
                          int idx_bytes, int64_t ro_s, int64_t co_s, const T* A, int64_t lda, T beta, T* B, int64_t ldb,
                          void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(ok_layout(layout) && ok_op(opS) && ok_op(opA));
    RB_REQUIRE(d >= 0 && n >= 0 && m >= 0 && ro_s >= 0 && co_s >= 0 && nnz >= 0);
    RB_REQUIRE(idx_bytes == 4 || idx_bytes == 8);
    const int64_t a1 = left ? d : n, a2 = left ? m : d;
    const int64_t rs = (opS == 'N') ? a1 : a2, cs = (opS == 'N') ? a2 : a1;
    RB_REQUIRE(S_rows >= rs + ro_s);
    RB_REQUIRE(S_cols >= cs + co_s);
    const int64_t rows_A = (opA == 'N') ? m : n, cols_A = (opA == 'N') ? n : m;
    const int64_t rows_B = left ? d : m, cols_B = left ? n : d;
    if (layout == 'C') { RB_REQUIRE(lda >= rows_A); RB_REQUIRE(ldb >= rows_B); }
    else               { RB_REQUIRE(lda >= cols_A); RB_REQUIRE(ldb >= cols_B); }
    if (rows_B == 0 || cols_B == 0) return 0;
    RB_REQUIRE(B != nullptr);
    Staged sA, sB, sv, sr, sc;
    const int64_t A_outer = (layout == 'C') ? cols_A : rows_A, A_inner = (layout == 'C') ? rows_A : cols_A;
    const int64_t B_outer = (layout == 'C') ? cols_B : rows_B, B_inner = (layout == 'C') ? rows_B : cols_B;
    int rc = sA.open(A, sizeof(T), A_outer, A_inner, lda, true, false, st); if (rc) return rc;
    rc = sB.open(B, sizeof(T), B_outer, B_inner, ldb, beta != (T) 0, true, st); if (rc) return rc;
    rc = sv.open(vals, sizeof(T), 1, nnz, nnz, true, false, st); if (rc) return rc;
    rc = sr.open(rows, (size_t) idx_bytes, 1, nnz, nnz, true, false, st); if (rc) return rc;
    rc = sc.open(cols, (size_t) idx_bytes, 1, nnz, nnz, true, false, st); if (rc) return rc;
    const int64_t ars = (layout == 'C') ? 1 : lda, acs = (layout == 'C') ? lda : 1;
    const int64_t brs = (layout == 'C') ? 1 : ldb, bcs = (layout == 'C') ? ldb : 1;
    CooProblem<T> c;
    c.alpha = alpha; c.beta = beta; c.nnz = nnz; c.vals = (const T*) sv.dev; c.rows = sr.dev; c.cols = sc.dev;
    c.idx_bytes = idx_bytes; c.ro_s = ro_s; c.co_s = co_s; c.rs = rs; c.cs = cs;
    c.Y = (const T*) sA.dev; c.C = (T*) sB.dev;
    if (left) {
        c.P = d; c.Q = n; c.K = m; c.x_is_transposed = (opS == 'T');
        c.yrs = (opA == 'N') ? ars : acs; c.ycs = (opA == 'N') ? acs : ars; c.crs = brs; c.ccs = bcs;
    } else {
        c.P = d; c.Q = m; c.K = n; c.x_is_transposed = (opS == 'N');
        c.yrs = (opA == 'N') ? acs : ars; c.ycs = (opA == 'N') ? ars : acs; c.crs = bcs; c.ccs = brs;
    }
    rc = launch_coo_apply<T>(c, st);
    int rc2;
    rc2 = sv.close(); if (!rc) rc = rc2;
    rc2 = sr.close(); if (!rc) rc = rc2;
    rc2 = sc.close(); if (!rc) rc = rc2;
    rc2 = sA.close(); if (!rc) rc = rc2;
    rc2 = sB.close(); if (!rc) rc = rc2;
    return rc;
}

// util::require_symmetric [util.hh:128-148], the check in front of sketch_symmetric [sksy.hh:159-176, 294-312]
template <typename T>
static int require_symmetric_impl(char layout, const T* A, int64_t n, int64_t lda, T tol, void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    if (tol < (T) 0) return 0;
    RB_REQUIRE(ok_layout(layout));
    RB_REQUIRE(n >= 0 && lda >= n);
    if (n <= 1) return 0;
    RB_REQUIRE(A != nullptr);
    Staged sA;
    int rc = sA.open(A, sizeof(T), n, n, lda, true, false, st); if (rc) return rc;
    unsigned long long* first = (unsigned long long*) workspace(2, 8, st);
    if (!first) return fail_cuda(cudaErrorMemoryAllocation, "symmetry check workspace");
    const int64_t rs = (layout == 'C') ? 1 : lda, cs = (layout == 'C') ? lda : 1;
    rc = launch_symmetry_check<T>((const T*) sA.dev, n, rs, cs, tol, first, st);
    unsigned long long h = ~0ull;
    if (!rc) {
        if (cudaMemcpyAsync(&h, first, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
            rc = fail_cuda(cudaGetLastError(), "symmetry check");
    }
    int rc2 = sA.close(); if (!rc) rc = rc2;
    if (rc) return rc;
    if (h != ~0ull) {
        const long long i = (long long) (h / (unsigned long long) n), j = (long long) (h % (unsigned long long) n);
        char msg[200];
        std::snprintf(msg, sizeof msg, "Symmetry check failed. |A(%lld,%lld) - A(%lld,%lld)| exceeds the tolerance of %e.", i, j,
                      j, i, (double) tol);
        return fail(msg);
    }
    return 0;
}

// sparse data matrix applied to a dense matrix: left_spmm / right_spmm [spmm_dispatch.hh:52-219].
//   left : C(d x n) = alpha * op(A_sp[ro_a:, co_a:])(d x m) * op(B)(m x n) + beta * C
//   right: C(m x d) = alpha * op(B)(m x n) * op(A_sp[ro_a:, co_a:])(n x d) + beta * C
// COO goes to the COO kernel as is; CSR / CSC first get their pointer array expanded to an index array.
template <typename T>
static int spmm_impl(int left, int fmt, char layout, char opA, char opB, int64_t d, int64_t n, int64_t m, T alpha,
                     int64_t A_rows, int64_t A_cols, int64_t nnz, const T* vals, const void* idx0, const void* idx1,
                     int idx_bytes, int64_t ro_a, int64_t co_a, const T* B, int64_t ldb, T beta, T* C, int64_t ldc,
                     void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(fmt >= 0 && fmt <= 2);
    RB_REQUIRE(ok_op(opA));
    RB_REQUIRE(nnz >= 0 && A_rows >= 0 && A_cols >= 0);
    RB_REQUIRE(idx_bytes == 4 || idx_bytes == 8);
    if (fmt != 2) {
        // spmm_dispatch.hh:99-107 (after the transposition of :87-88): compressed formats take no submatrix
        const int64_t a1 = left ? d : n, a2 = left ? m : d;              // op(A_sp window) is a1 x a2
        RB_REQUIRE(((opA == 'N') ? A_rows : A_cols) == a1);
        RB_REQUIRE(((opA == 'N') ? A_cols : A_rows) == a2);
        RB_REQUIRE(ro_a == 0);
        RB_REQUIRE(co_a == 0);
    }
    if (fmt == 2)
        return coo_apply_impl<T>(left, layout, opA, opB, d, n, m, alpha, A_rows, A_cols, nnz, vals, idx0, idx1, idx_bytes,
                                 ro_a, co_a, B, ldb, beta, C, ldc, stream);
    // CSR: idx0 = rowptr (A_rows + 1), idx1 = colidxs; CSC: idx0 = rowidxs, idx1 = colptr (A_cols + 1)
    const int64_t n_major = (fmt == 0) ? A_rows : A_cols;
    const void* ptr = (fmt == 0) ? idx0 : idx1;
    RB_REQUIRE(ptr != nullptr || n_major == 0);
    Staged sp;
    int rc = sp.open(ptr, (size_t) idx_bytes, 1, n_major + 1, n_major + 1, true, false, st); if (rc) return rc;
    void* expanded = workspace(3, (size_t) (nnz > 0 ? nnz : 1) * (size_t) idx_bytes, st);
    if (!expanded) return fail_cuda(cudaErrorMemoryAllocation, "spmm index workspace");
    rc = launch_expand_ptr(n_major, sp.dev, expanded, idx_bytes, st);
    int rc2 = sp.close(); if (!rc) rc = rc2;
    if (rc) return rc;
    const void* rows = (fmt == 0) ? expanded : idx0;
    const void* cols = (fmt == 0) ? idx1 : expanded;
    return coo_apply_impl<T>(left, layout, opA, opB, d, n, m, alpha, A_rows, A_cols, nnz, vals, rows, cols, idx_bytes, ro_a,
                             co_a, B, ldb, beta, C, ldc, stream);
}

// dense operator applied to sparse data [sksp.hh:132-182, 277-326]
template <typename T>
static int sksp3_impl(bool left, int fmt, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, T alpha,
                      int64_t D_rows, int64_t D_cols, char family, char axis, const uint32_t* ctr, const uint32_t* key,
                      int64_t ro_s, int64_t co_s, int64_t A_rows, int64_t A_cols, int64_t nnz, const T* vals,
                      const void* idx0, const void* idx1, int idx_bytes, int64_t ro_a, int64_t co_a, T beta, T* B,
                      int64_t ldb, void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(fmt >= 0 && fmt <= 2);
    RB_REQUIRE(ok_layout(layout) && ok_op(opS) && ok_op(opA));
    RB_REQUIRE(d >= 0 && n >= 0 && m >= 0 && ro_s >= 0 && co_s >= 0 && ro_a >= 0 && co_a >= 0 && nnz >= 0);
    RB_REQUIRE(idx_bytes == 4 || idx_bytes == 8);
    DenseDistInfo D = make_dense_dist(D_rows, D_cols, family, axis);
    RB_REQUIRE(D.ok);
    RB_REQUIRE(ctr != nullptr && key != nullptr);
    const int64_t a1 = left ? d : n, a2 = left ? m : d;
    const int64_t rows_sub = (opS == 'N') ? a1 : a2, cols_sub = (opS == 'N') ? a2 : a1;
    const int64_t rows_subA = (opA == 'N') ? m : n, cols_subA = (opA == 'N') ? n : m;
    RB_REQUIRE(A_rows >= rows_subA + ro_a);        // sksp.hh:165, 309
    RB_REQUIRE(A_cols >= cols_subA + co_a);        // sksp.hh:166, 310
    RB_REQUIRE(D.n_rows >= rows_sub + ro_s);       // sksp.hh:167, 311
    RB_REQUIRE(D.n_cols >= cols_sub + co_s);       // sksp.hh:168, 312
    const int64_t rows_B = left ? d : m, cols_B = left ? n : d;
    if (layout == 'C') RB_REQUIRE(ldb >= rows_B); else RB_REQUIRE(ldb >= cols_B);   // sksp.hh:169-173, 313-317
    if (rows_B == 0 || cols_B == 0) return 0;
    RB_REQUIRE(B != nullptr);
    RB_REQUIRE(nnz == 0 || (vals != nullptr && idx0 != nullptr && idx1 != nullptr));

    Staged sB, sv, s0, s1;
    const int64_t B_outer = (layout == 'C') ? cols_B : rows_B, B_inner = (layout == 'C') ? rows_B : cols_B;
    const int64_t n0 = (fmt == 0) ? A_rows + 1 : nnz, n1 = (fmt == 1) ? A_cols + 1 : nnz;
    int rc = sB.open(B, sizeof(T), B_outer, B_inner, ldb, beta != (T) 0, true, st); if (rc) return rc;
    rc = sv.open(vals, sizeof(T), 1, nnz, nnz, true, false, st); if (rc) return rc;
    rc = s0.open(idx0, (size_t) idx_bytes, 1, n0, n0, true, false, st); if (rc) return rc;
    rc = s1.open(idx1, (size_t) idx_bytes, 1, n1, n1, true, false, st); if (rc) return rc;

    const int64_t brs = (layout == 'C') ? 1 : ldb, bcs = (layout == 'C') ? ldb : 1;
    SpDataProblem<T> p;
    p.alpha = alpha; p.beta = beta;
    p.gen = make_dense_gen(D, ctr, key);
    p.family = family;
    p.fmt = fmt; p.A_rows = A_rows; p.A_cols = A_cols; p.nnz = nnz;
    p.vals = (const T*) sv.dev; p.idx0 = s0.dev; p.idx1 = s1.dev; p.idx_bytes = idx_bytes;
    p.ro_a = ro_a; p.co_a = co_a;
    p.C = (T*) sB.dev;
    OpWindow w;
    if (left) {
        p.P = d; p.Q = n; p.K = m;
        w = op_window(p.gen.nat_row, opS, ro_s, co_s);
        p.y_is_transposed = (opA == 'T');
        p.crs = brs; p.ccs = bcs;
    } else {
        p.P = d; p.Q = m; p.K = n;
        w = op_window(p.gen.nat_row, opS == 'N' ? 'T' : 'N', ro_s, co_s);
        p.y_is_transposed = (opA == 'N');
        p.crs = bcs; p.ccs = brs;
    }
    p.v0 = w.v0; p.u0 = w.u0; p.vi = w.vi; p.ui = w.ui; p.vk = w.vk; p.uk = w.uk;
    rc = launch_spdata<T>(p, st);
    int rc2;
    rc2 = sv.close(); if (!rc) rc = rc2;
    rc2 = s0.close(); if (!rc) rc = rc2;
    rc2 = s1.close(); if (!rc) rc = rc2;
    rc2 = sB.close(); if (!rc) rc = rc2;
    return rc;
}

}  // namespace rb

using namespace rb;

// weights_to_cdf<T> (util.hh:459-473). The prefix written before a failing weight stays in w, as in the reference.
template <typename T>
static int weights_to_cdf_impl(int64_t n, T* w, T error_if_below, void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(n >= 0);
    if (n == 0) return 0;
    RB_REQUIRE(w != nullptr);
    Staged sw;
    int rc = sw.open(w, sizeof(T), 1, n, n, true, true, st); if (rc) return rc;
    rc = launch_weights_to_cdf<T>(n, (T*) sw.dev, error_if_below, st);
    int rc2 = sw.close(); if (!rc) rc = rc2;
    return rc;
}

extern "C" {

const char* rb_last_error(void) { return g_err.c_str(); }
int rb_version(void) { return 100; }
int rb_release_workspace(void) {
    release_workspace();
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
        cudaMemPoolTrimTo(pool, 0);            // staging buffers kept by keep_pool_memory()
    cudaGetLastError();
    return 0;
}

int rb_device_info(int64_t info[3]) {
    int dev = 0;
    RB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    RB_CUDA(cudaGetDeviceProperties(&prop, dev));
    info[0] = prop.multiProcessorCount;
    info[1] = prop.major * 10 + prop.minor;
    info[2] = (int64_t) prop.totalGlobalMem;
    return 0;
}

int rb_sync_stream(void* stream) {
    RB_CUDA(cudaStreamSynchronize((cudaStream_t) stream));
    return 0;
}

void rb_rngstate_from_u64(uint64_t k, uint32_t ctr[4], uint32_t key[2]) {
    ctr[0] = ctr[1] = ctr[2] = ctr[3] = 0;
    key[0] = (uint32_t) k;
    key[1] = (uint32_t) (k >> 32);
}
void rb_ctr_incr(uint32_t ctr[4], uint64_t n) { store_ctr(ctr_add(load_ctr(ctr), n), ctr); }

int rb_dense_dist_info(int64_t n_rows, int64_t n_cols, char family, char major_axis, int64_t info[3],
                       double* isometry_scale) {
    DenseDistInfo D = make_dense_dist(n_rows, n_cols, family, major_axis);
    RB_REQUIRE(D.ok);
    info[0] = D.dim_major; info[1] = D.dim_minor; info[2] = D.natural_layout;
    if (isometry_scale) *isometry_scale = std::pow((double) D.dim_minor, -0.5);      // dense_skops.hh:322
    return 0;
}
int rb_sparse_dist_info(int64_t n_rows, int64_t n_cols, int64_t vec_nnz, char major_axis, int64_t info[3],
                        double* isometry_scale) {
    SparseDistInfo D = make_sparse_dist(n_rows, n_cols, vec_nnz, major_axis);
    RB_REQUIRE(D.ok);
    info[0] = D.dim_major; info[1] = D.dim_minor; info[2] = D.full_nnz;
    if (isometry_scale) {                                                              // sparse_skops.hh:108-114
        if (major_axis == 'S') *isometry_scale = std::pow((double) vec_nnz, -0.5);
        else *isometry_scale = std::sqrt(((double) D.dim_major) / (vec_nnz * ((double) D.dim_minor)));
    }
    return 0;
}
int rb_dense_next_state(int64_t n_rows, int64_t n_cols, char family, char major_axis, const uint32_t ctr[4],
                        uint32_t next_ctr[4]) {
    DenseDistInfo D = make_dense_dist(n_rows, n_cols, family, major_axis);
    RB_REQUIRE(D.ok);
    store_ctr(ctr_add(load_ctr(ctr), (uint64_t) (((D.dim_major + 3) / 4) * D.dim_minor)), next_ctr);
    return 0;
}
int rb_sparse_next_state(int64_t n_rows, int64_t n_cols, int64_t vec_nnz, char major_axis, const uint32_t ctr[4],
                         uint32_t next_ctr[4]) {
    SparseDistInfo D = make_sparse_dist(n_rows, n_cols, vec_nnz, major_axis);
    RB_REQUIRE(D.ok);
    int64_t num_mavec, incrs;
    if (major_axis == 'S') { num_mavec = std::max(n_rows, n_cols); incrs = vec_nnz; }
    else { num_mavec = std::min(n_rows, n_cols); incrs = (int64_t) std::ceil((double) vec_nnz / 2.0); }
    store_ctr(ctr_add(load_ctr(ctr), (uint64_t) (num_mavec * incrs)), next_ctr);
    return 0;
}

// coo_to_csr / coo_to_csc [sparse_data/conversions.hh:79-121]; to_csc = 0: rows are compressed, 1: columns
int rb_coo_to_compressed(int to_csc, int64_t n_rows, int64_t n_cols, int64_t nnz, const void* vals, int val_bytes,
                         const void* rows, const void* cols, int idx_bytes, void* out_vals, void* out_idx, void* out_ptr,
                         void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0);
    RB_REQUIRE(idx_bytes == 4 || idx_bytes == 8);
    RB_REQUIRE(val_bytes == 4 || val_bytes == 8);
    RB_REQUIRE(out_ptr != nullptr);
    RB_REQUIRE(nnz == 0 || (vals && rows && cols && out_vals && out_idx));
    const int64_t n_major = to_csc ? n_cols : n_rows, n_minor = to_csc ? n_rows : n_cols;
    Staged sv, sr, sc, ov, oi, op;
    int rc = sv.open(vals, (size_t) val_bytes, 1, nnz, nnz, true, false, st); if (rc) return rc;
    rc = sr.open(rows, (size_t) idx_bytes, 1, nnz, nnz, true, false, st); if (rc) return rc;
    rc = sc.open(cols, (size_t) idx_bytes, 1, nnz, nnz, true, false, st); if (rc) return rc;
    rc = ov.open(out_vals, (size_t) val_bytes, 1, nnz, nnz, false, true, st); if (rc) return rc;
    rc = oi.open(out_idx, (size_t) idx_bytes, 1, nnz, nnz, false, true, st); if (rc) return rc;
    rc = op.open(out_ptr, (size_t) idx_bytes, 1, n_major + 1, n_major + 1, false, true, st); if (rc) return rc;
    rc = launch_coo_to_compressed(n_major, n_minor, nnz, sv.dev, val_bytes, to_csc ? sc.dev : sr.dev, to_csc ? sr.dev : sc.dev,
                                  idx_bytes, ov.dev, oi.dev, op.dev, nullptr, st);
    int rc2 = sv.close(); if (!rc) rc = rc2;
    rc2 = sr.close(); if (!rc) rc = rc2;
    rc2 = sc.close(); if (!rc) rc = rc2;
    rc2 = ov.close(); if (!rc) rc = rc2;
    rc2 = oi.close(); if (!rc) rc = rc2;
    rc2 = op.close(); if (!rc) rc = rc2;
    return rc;
}

// csr_to_coo / csc_to_coo [sparse_data/conversions.hh:49-75]: the pointer array expanded to one index per entry
int rb_expand_ptr(int64_t n_major, const void* ptr, int64_t nnz, void* out_idx, int idx_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(n_major >= 0 && nnz >= 0);
    RB_REQUIRE(idx_bytes == 4 || idx_bytes == 8);
    RB_REQUIRE(ptr != nullptr);
    RB_REQUIRE(nnz == 0 || out_idx != nullptr);
    Staged sp, so;
    int rc = sp.open(ptr, (size_t) idx_bytes, 1, n_major + 1, n_major + 1, true, false, st); if (rc) return rc;
    rc = so.open(out_idx, (size_t) idx_bytes, 1, nnz, nnz, false, true, st); if (rc) return rc;
    rc = launch_expand_ptr(n_major, sp.dev, so.dev, idx_bytes, st);
    int rc2 = sp.close(); if (!rc) rc = rc2;
    rc2 = so.close(); if (!rc) rc = rc2;
    return rc;
}

int rb_philox_words(const uint32_t ctr[4], const uint32_t key[2], int64_t n_blocks, uint32_t* out, void* stream) {
    RB_REQUIRE(ctr != nullptr && key != nullptr && n_blocks >= 0);
    if (n_blocks == 0) return 0;
    RB_REQUIRE(out != nullptr);
    Staged so;
    int rc = so.open(out, 16, 1, n_blocks, n_blocks, false, true, (cudaStream_t) stream);
    if (rc) return rc;
    rc = launch_philox_words(load_ctr(ctr), PhiloxKey{key[0], key[1]}, n_blocks, (uint32_t*) so.dev, (cudaStream_t) stream);
    int rc2 = so.close();
    return rc ? rc : rc2;
}

int rb_boxmuller_words(int64_t n, const uint32_t* w0, const uint32_t* w1, float* g0, float* g1, void* stream) {
    RB_REQUIRE(n >= 0);
    if (n == 0) return 0;
    RB_REQUIRE(w0 != nullptr && w1 != nullptr && g0 != nullptr && g1 != nullptr);
    cudaStream_t st = (cudaStream_t) stream;
    Staged s0, s1, o0, o1;
    int rc = s0.open(w0, 4, 1, n, n, true, false, st); if (rc) return rc;
    rc = s1.open(w1, 4, 1, n, n, true, false, st); if (rc) return rc;
    rc = o0.open(g0, 4, 1, n, n, false, true, st); if (rc) return rc;
    rc = o1.open(g1, 4, 1, n, n, false, true, st); if (rc) return rc;
    rc = launch_boxmuller_words(n, (const uint32_t*) s0.dev, (const uint32_t*) s1.dev, (float*) o0.dev, (float*) o1.dev, st);
    int rc2;
    rc2 = s0.close(); if (!rc) rc = rc2;
    rc2 = s1.close(); if (!rc) rc = rc2;
    rc2 = o0.close(); if (!rc) rc = rc2;
    rc2 = o1.close(); if (!rc) rc = rc2;
    return rc;
}

int rb_fill_dense_f32(char layout, int64_t D_rows, int64_t D_cols, char family, char major_axis, int64_t n_rows,
                      int64_t n_cols, int64_t ro_s, int64_t co_s, float* buff, int64_t ld, const uint32_t ctr[4],
                      const uint32_t key[2], uint32_t next_ctr[4], void* stream) {
    return fill_dense_impl<float>(layout, D_rows, D_cols, family, major_axis, n_rows, n_cols, ro_s, co_s, buff, ld, ctr,
                                  key, next_ctr, stream);
}
int rb_fill_dense_f64(char layout, int64_t D_rows, int64_t D_cols, char family, char major_axis, int64_t n_rows,
                      int64_t n_cols, int64_t ro_s, int64_t co_s, double* buff, int64_t ld, const uint32_t ctr[4],
                      const uint32_t key[2], uint32_t next_ctr[4], void* stream) {
    return fill_dense_impl<double>(layout, D_rows, D_cols, family, major_axis, n_rows, n_cols, ro_s, co_s, buff, ld,
                                   ctr, key, next_ctr, stream);
}

int rb_fill_sparse_saso(int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4], const uint32_t key[2],
                        void* vals, int val_bytes, void* rows, void* cols, int idx_bytes, int64_t* nnz,
                        uint32_t next_ctr[4], void* stream) {
    return fill_sparse_impl(D_rows, D_cols, vec_nnz, ctr, key, vals, val_bytes, rows, cols, idx_bytes, nnz, next_ctr,
                            stream);
}

int rb_fill_sparse_laso(int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4], const uint32_t key[2],
                        void* vals, int val_bytes, void* rows, void* cols, int idx_bytes, int64_t* nnz,
                        uint32_t next_ctr[4], void* stream) {
    return fill_sparse_laso_impl(D_rows, D_cols, vec_nnz, ctr, key, vals, val_bytes, rows, cols, idx_bytes, nnz, next_ctr,
                                 stream);
}

int rb_repeated_fisher_yates(int64_t k, int64_t n, int64_t r, void* samples, int idx_bytes, const uint32_t ctr[4],
                             const uint32_t key[2], uint32_t next_ctr[4], void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(k >= 0 && n >= 0 && r >= 0);
    RB_REQUIRE(k <= n);                            // sparse_skops.hh:63
    RB_REQUIRE(idx_bytes == 4 || idx_bytes == 8);
    RB_REQUIRE(ctr != nullptr && key != nullptr);
    const Ctr128 c = load_ctr(ctr);
    store_ctr(ctr_add(c, (uint64_t) (k * r)), next_ctr);
    if (k == 0 || r == 0) return 0;
    RB_REQUIRE(samples != nullptr);
    Staged ss;
    int rc = ss.open(samples, (size_t) idx_bytes, 1, k * r, k * r, false, true, st);
    if (rc) return rc;
    rc = launch_saso(c, PhiloxKey{key[0], key[1]}, k, n, r, ss.dev, nullptr, idx_bytes, nullptr, 4, st);
    int rc2 = ss.close();
    return rc ? rc : rc2;
}

// sample_indices_iid_uniform<T, sint_t, WriteRademachers> (util.hh:515-560)
int rb_sample_indices_iid_uniform(int64_t n, int64_t k, void* samples, int idx_bytes, void* rademachers, int val_bytes,
                                  const uint32_t ctr[4], const uint32_t key[2], uint32_t next_ctr[4], void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(n >= 0 && k >= 0);
    RB_REQUIRE(idx_bytes == 4 || idx_bytes == 8);
    RB_REQUIRE(rademachers == nullptr || val_bytes == 4 || val_bytes == 8);
    RB_REQUIRE(ctr != nullptr && key != nullptr);
    if (idx_bytes == 4) RB_REQUIRE(n <= 2147483647LL);
    const Ctr128 c = load_ctr(ctr);
    const int64_t spb = rademachers ? 2 : 4;                       // util.hh:521-525
    store_ctr(ctr_add(c, (uint64_t) ((k + spb - 1) / spb)), next_ctr);   // one incr per block used, util.hh:538-544
    if (k == 0) return 0;
    RB_REQUIRE(samples != nullptr);
    Staged ss, sr;
    int rc = ss.open(samples, (size_t) idx_bytes, 1, k, k, false, true, st); if (rc) return rc;
    rc = sr.open(rademachers, (size_t) (rademachers ? val_bytes : 4), 1, k, k, false, true, st); if (rc) return rc;
    rc = launch_sample_indices_iid_uniform(c, PhiloxKey{key[0], key[1]}, n, k, ss.dev, idx_bytes, sr.dev, val_bytes, st);
    int rc2 = ss.close(); if (!rc) rc = rc2;
    rc2 = sr.close(); if (!rc) rc = rc2;
    return rc;
}

// sample_indices_iid<T, sint_t> (util.hh:490-513)
int rb_sample_indices_iid(int64_t n, const void* cdf, int val_bytes, int64_t k, void* samples, int idx_bytes,
                          const uint32_t ctr[4], const uint32_t key[2], uint32_t next_ctr[4], void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(n >= 0 && k >= 0);
    RB_REQUIRE(idx_bytes == 4 || idx_bytes == 8);
    RB_REQUIRE(val_bytes == 4 || val_bytes == 8);
    RB_REQUIRE(ctr != nullptr && key != nullptr);
    if (idx_bytes == 4) RB_REQUIRE(n <= 2147483647LL);
    const Ctr128 c = load_ctr(ctr);
    store_ctr(ctr_add(c, (uint64_t) ((k + 3) / 4)), next_ctr);     // util.hh:505-511
    if (k == 0) return 0;
    RB_REQUIRE(samples != nullptr);
    RB_REQUIRE(n == 0 || cdf != nullptr);
    Staged sc, ss;
    int rc = sc.open(cdf, (size_t) val_bytes, 1, n, n, true, false, st); if (rc) return rc;
    rc = ss.open(samples, (size_t) idx_bytes, 1, k, k, false, true, st); if (rc) return rc;
    rc = launch_sample_indices_iid(c, PhiloxKey{key[0], key[1]}, n, sc.dev, val_bytes, k, ss.dev, idx_bytes, st);
    int rc2 = ss.close(); if (!rc) rc = rc2;
    rc2 = sc.close(); if (!rc) rc = rc2;
    return rc;
}

int rb_weights_to_cdf_f32(int64_t n, float* w, float error_if_below, void* stream) {
    return weights_to_cdf_impl<float>(n, w, error_if_below, stream);
}
int rb_weights_to_cdf_f64(int64_t n, double* w, double error_if_below, void* stream) {
    return weights_to_cdf_impl<double>(n, w, error_if_below, stream);
}

#define RB_DEF_T(T, sfx)                                                                                               \
    int rb_lskge3_##sfx(char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, T alpha, int64_t D_rows,     \
                        int64_t D_cols, char family, char major_axis, const uint32_t ctr[4], const uint32_t key[2],    \
                        const T* S_buff, int64_t ro_s, int64_t co_s, const T* A, int64_t lda, T beta, T* B,            \
                        int64_t ldb, void* stream) {                                                                   \
        return skge3_impl<T>(true, layout, opS, opA, d, n, m, alpha, D_rows, D_cols, family, major_axis, ctr, key,     \
                             S_buff, ro_s, co_s, A, lda, beta, B, ldb, stream);                                        \
    }                                                                                                                  \
    int rb_rskge3_##sfx(char layout, char opA, char opS, int64_t m, int64_t d, int64_t n, T alpha, const T* A,         \
                        int64_t lda, int64_t D_rows, int64_t D_cols, char family, char major_axis,                     \
                        const uint32_t ctr[4], const uint32_t key[2], const T* S_buff, int64_t ro_s, int64_t co_s,     \
                        T beta, T* B, int64_t ldb, void* stream) {                                                     \
        return skge3_impl<T>(false, layout, opS, opA, d, n, m, alpha, D_rows, D_cols, family, major_axis, ctr, key,    \
                             S_buff, ro_s, co_s, A, lda, beta, B, ldb, stream);                                        \
    }                                                                                                                  \
    int rb_lskges_##sfx(char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, T alpha, int64_t D_rows,     \
                        int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s,   \
                        int64_t co_s, const T* A, int64_t lda, T beta, T* B, int64_t ldb, void* stream) {              \
        return skges_impl<T>(true, layout, opS, opA, d, n, m, alpha, D_rows, D_cols, vec_nnz, ctr, key, ro_s, co_s, A, \
                             lda, beta, B, ldb, stream);                                                               \
    }                                                                                                                  \
    int rb_rskges_##sfx(char layout, char opA, char opS, int64_t m, int64_t d, int64_t n, T alpha, const T* A,         \
                        int64_t lda, int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4],           \
                        const uint32_t key[2], int64_t ro_s, int64_t co_s, T beta, T* B, int64_t ldb, void* stream) {  \
        return skges_impl<T>(false, layout, opS, opA, d, n, m, alpha, D_rows, D_cols, vec_nnz, ctr, key, ro_s, co_s,   \
                             A, lda, beta, B, ldb, stream);                                                            \
    }                                                                                                                  \
    int rb_coo_apply_##sfx(int side_left, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, T alpha,   \
                           int64_t S_rows, int64_t S_cols, int64_t nnz, const T* vals, const void* rows,               \
                           const void* cols, int idx_bytes, int64_t ro_s, int64_t co_s, const T* A, int64_t lda,       \
                           T beta, T* B, int64_t ldb, void* stream) {                                                  \
        return coo_apply_impl<T>(side_left, layout, opS, opA, d, n, m, alpha, S_rows, S_cols, nnz, vals, rows, cols,   \
                                 idx_bytes, ro_s, co_s, A, lda, beta, B, ldb, stream);                                 \
    }                                                                                                                  \
    int rb_require_symmetric_##sfx(char layout, const T* A, int64_t n, int64_t lda, T tol, void* stream) {             \
        return require_symmetric_impl<T>(layout, A, n, lda, tol, stream);                                              \
    }                                                                                                                  \
    int rb_spmm_##sfx(int side_left, int fmt, char layout, char opA, char opB, int64_t d, int64_t n, int64_t m,         \
                      T alpha, int64_t A_rows, int64_t A_cols, int64_t nnz, const T* vals, const void* idx0,           \
                      const void* idx1, int idx_bytes, int64_t ro_a, int64_t co_a, const T* B, int64_t ldb, T beta,    \
                      T* C, int64_t ldc, void* stream) {                                                               \
        return spmm_impl<T>(side_left, fmt, layout, opA, opB, d, n, m, alpha, A_rows, A_cols, nnz, vals, idx0, idx1,   \
                            idx_bytes, ro_a, co_a, B, ldb, beta, C, ldc, stream);                                      \
    }                                                                                                                  \
    int rb_lsksp3_##sfx(int fmt, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, T alpha,            \
                        int64_t D_rows, int64_t D_cols, char family, char major_axis, const uint32_t ctr[4],           \
                        const uint32_t key[2], int64_t ro_s, int64_t co_s, int64_t A_rows, int64_t A_cols,             \
                        int64_t nnz, const T* vals, const void* idx0, const void* idx1, int idx_bytes, int64_t ro_a,   \
                        int64_t co_a, T beta, T* B, int64_t ldb, void* stream) {                                       \
        return sksp3_impl<T>(true, fmt, layout, opS, opA, d, n, m, alpha, D_rows, D_cols, family, major_axis, ctr,     \
                             key, ro_s, co_s, A_rows, A_cols, nnz, vals, idx0, idx1, idx_bytes, ro_a, co_a, beta, B,   \
                             ldb, stream);                                                                             \
    }                                                                                                                  \
    int rb_rsksp3_##sfx(int fmt, char layout, char opA, char opS, int64_t m, int64_t d, int64_t n, T alpha,            \
                        int64_t A_rows, int64_t A_cols, int64_t nnz, const T* vals, const void* idx0,                  \
                        const void* idx1, int idx_bytes, int64_t ro_a, int64_t co_a, int64_t D_rows, int64_t D_cols,   \
                        char family, char major_axis, const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s,      \
                        int64_t co_s, T beta, T* B, int64_t ldb, void* stream) {                                       \
        return sksp3_impl<T>(false, fmt, layout, opS, opA, d, n, m, alpha, D_rows, D_cols, family, major_axis, ctr,    \
                             key, ro_s, co_s, A_rows, A_cols, nnz, vals, idx0, idx1, idx_bytes, ro_a, co_a, beta, B,   \
                             ldb, stream);                                                                             \
    }

RB_DEF_T(float, f32)
RB_DEF_T(double, f64)

int64_t rb_get_option(const char* name) { return name ? rb::get_option(name) : 0; }

int rb_set_option(const char* name, int64_t value) {
    RB_REQUIRE(name != nullptr);
    if (!std::strcmp(name, "dense_path")) { g_dense_path = value; return 0; }
    if (!std::strcmp(name, "saso_fill_path")) { g_saso_fill_path = value; return 0; }
    if (!std::strcmp(name, "tc_splits")) { g_tc_splits = value; return 0; }
    if (!std::strcmp(name, "tc_cluster")) { g_tc_cluster = value; return 0; }
    if (!std::strcmp(name, "tc_pair")) { g_tc_pair = value; return 0; }
    if (!std::strcmp(name, "tc_ymn")) { g_tc_ymn = value; return 0; }
    if (!std::strcmp(name, "tc_xmn")) { g_tc_xmn = value; return 0; }
    if (!std::strcmp(name, "fill_rep")) { g_fill_rep = value; return 0; }
    if (!std::strcmp(name, "fill_unroll")) { g_fill_unroll = value; return 0; }
    if (!std::strcmp(name, "saso_rows")) { g_saso_rows = value; return 0; }
    if (!std::strcmp(name, "saso_bin_path")) { g_saso_bin_path = value; return 0; }
    if (!std::strcmp(name, "dmma_materialise")) { g_dmma_materialise = value; return 0; }
    if (!std::strcmp(name, "dmma_panel_mb")) { g_dmma_panel_mb = value; return 0; }
    if (!std::strcmp(name, "tc_materialise")) { g_tc_materialise = value; return 0; }
    if (!std::strcmp(name, "tc_halves")) { g_tc_halves = value; return 0; }
    if (!std::strcmp(name, "saso_path")) { g_saso_path = value; return 0; }
    if (!std::strcmp(name, "h2d_chunk_mb")) { RB_REQUIRE(value >= 1 && value <= 4096); g_h2d_chunk_mb = value; return 0; }
    if (!std::strcmp(name, "spdata_path")) { g_spdata_path = value; return 0; }
    return fail(std::string("unknown option ") + name);
}
int64_t rb_get_counter(const char* name) {
    if (!name) return -1;
    if (!std::strcmp(name, "kernel_launches")) return g_launches.load();
    if (!std::strcmp(name, "tensor_core_launches")) return g_tc_launches.load();
    if (!std::strcmp(name, "saso_owner_launches")) return g_owner_launches.load();
    return -1;
}

}  // extern "C"
