// Multi-GPU entry points of the C ABI: the m-sharded left sketch with its one exchange step.
//
// Reference semantics being distributed: a left sketch whose contraction dimension is split into row blocks of A,
// block g using the operator columns [start_g, start_g + count_g) through the (ro_s, co_s) submatrix arguments, the
// block products summed into one B -- RandBLAS/skge.hh:174-181 and rtd/source/tutorial/sketch_updates.rst:198-213
// (the reference's own blocked form with beta = 1). Here block g lives on GPU g and the sum is one NCCL
// reduce-scatter (or all-reduce) of the d x n partial products over NVLink.
//
// NCCL is bound at run time with dlopen("libnccl.so.2"): in a process that already loaded torch's bundled NCCL the
// same library instance is used; a plain C++ caller gets the system library. The rest of the library does not
// depend on NCCL, and these entry points fail with RB_ERR_CUDA + a message if it cannot be loaded.
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/randblas_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace rb {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

static NcclApi* nccl_api() {
    static std::mutex mu;
    static NcclApi api;
    static bool tried = false;
    std::lock_guard<std::mutex> lk(mu);
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (h) {
            api.GetUniqueId = (decltype(api.GetUniqueId)) dlsym(h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank)) dlsym(h, "ncclCommInitRank");
            api.CommInitAll = (decltype(api.CommInitAll)) dlsym(h, "ncclCommInitAll");
            api.CommDestroy = (decltype(api.CommDestroy)) dlsym(h, "ncclCommDestroy");
            api.ReduceScatter = (decltype(api.ReduceScatter)) dlsym(h, "ncclReduceScatter");
            api.AllReduce = (decltype(api.AllReduce)) dlsym(h, "ncclAllReduce");
            api.GroupStart = (decltype(api.GroupStart)) dlsym(h, "ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd)) dlsym(h, "ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString)) dlsym(h, "ncclGetErrorString");
            api.GetVersion = (decltype(api.GetVersion)) dlsym(h, "ncclGetVersion");
            if (api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.ReduceScatter &&
                api.AllReduce && api.GroupStart && api.GroupEnd && api.GetErrorString)
                api.handle = h;
        }
    }
    return api.handle ? &api : nullptr;
}

static int fail_nccl(NcclApi* n, ncclResult_t r, const char* what) {
    set_error(std::string("NCCL error: ") + (n ? n->GetErrorString(r) : "library not loaded") + " in " + what);
    return RB_ERR_CUDA;
}
#define RB_NCCL(n, call)                                                  \
    do {                                                                  \
        ncclResult_t r_ = (call);                                         \
        if (r_ != ncclSuccess) return fail_nccl(n, r_, #call);            \
    } while (0)

}  // namespace rb

// The communicator handle: one per (process, GPU). Owns the buffer that holds this GPU's d x n partial product.
struct rb_comm {
    ncclComm_t nccl = nullptr;
    int nranks = 1, rank = 0, device = 0;
    void* partial = nullptr;
    size_t partial_bytes = 0;
    bool owns_nccl = true;
};

namespace rb {

static void* comm_partial(rb_comm* c, size_t bytes, cudaStream_t st) {
    if (c->partial_bytes < bytes) {
        if (c->partial) { cudaStreamSynchronize(st); cudaFree(c->partial); c->partial = nullptr; c->partial_bytes = 0; }
        if (cudaMalloc(&c->partial, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        c->partial_bytes = bytes;
    }
    return c->partial;
}

template <typename T>
__global__ void axpby_kernel(int64_t n, T beta, T* __restrict__ y, const T* __restrict__ x) {
    for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
        y[i] = beta * y[i] + x[i];
}

// rows of op(A) owned by `rank`: contiguous, starts multiples of 4 so that no Philox block of the operator is split
static void mshard_block(int64_t total, int rank, int world, int64_t* start, int64_t* count) {
    const int64_t units = (total + 3) / 4, per = units / world, rem = units % world;
    const int64_t u0 = rank * per + (rank < rem ? rank : rem), u1 = u0 + per + (rank < rem ? 1 : 0);
    const int64_t a = (u0 * 4 < total) ? u0 * 4 : total, b = (u1 * 4 < total) ? u1 * 4 : total;
    *start = a; *count = b - a;
}

template <typename T> struct Skge;
template <> struct Skge<float> {
    static constexpr ncclDataType_t nt = ncclFloat32;
    static int left(char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, float alpha, int64_t Dr, int64_t Dc,
                    char fam, char ax, const uint32_t* ctr, const uint32_t* key, int64_t ro, int64_t co, const float* A,
                    int64_t lda, float beta, float* B, int64_t ldb, void* st) {
        return rb_lskge3_f32(layout, opS, opA, d, n, m, alpha, Dr, Dc, fam, ax, ctr, key, nullptr, ro, co, A, lda, beta, B, ldb, st);
    }
    static int left_saso(char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, float alpha, int64_t Dr, int64_t Dc,
                         int64_t vec_nnz, const uint32_t* ctr, const uint32_t* key, int64_t ro, int64_t co, const float* A,
                         int64_t lda, float beta, float* B, int64_t ldb, void* st) {
        return rb_lskges_f32(layout, opS, opA, d, n, m, alpha, Dr, Dc, vec_nnz, ctr, key, ro, co, A, lda, beta, B, ldb, st);
    }
};
template <> struct Skge<double> {
    static constexpr ncclDataType_t nt = ncclFloat64;
    static int left(char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, double alpha, int64_t Dr, int64_t Dc,
                    char fam, char ax, const uint32_t* ctr, const uint32_t* key, int64_t ro, int64_t co, const double* A,
                    int64_t lda, double beta, double* B, int64_t ldb, void* st) {
        return rb_lskge3_f64(layout, opS, opA, d, n, m, alpha, Dr, Dc, fam, ax, ctr, key, nullptr, ro, co, A, lda, beta, B, ldb, st);
    }
    static int left_saso(char layout, char opS, char opA, int64_t d, int64_t n, int64_t m, double alpha, int64_t Dr, int64_t Dc,
                         int64_t vec_nnz, const uint32_t* ctr, const uint32_t* key, int64_t ro, int64_t co, const double* A,
                         int64_t lda, double beta, double* B, int64_t ldb, void* st) {
        return rb_lskges_f64(layout, opS, opA, d, n, m, alpha, Dr, Dc, vec_nnz, ctr, key, ro, co, A, lda, beta, B, ldb, st);
    }
};

// The operator of a sharded sketch: a DenseDist (family, axis) or, with vec_nnz > 0, a SASO SparseDist(D_rows, D_cols, vec_nnz).
struct ShardOp {
    int64_t D_rows, D_cols;
    char family, axis;
    int64_t vec_nnz;          // 0: dense operator
    const uint32_t *ctr, *key;
};

// Phase 1 of the sharded sketch on ONE GPU: this rank's partial product into the communicator's buffer.
template <typename T>
static int mshard_local(rb_comm* c, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m_total, T alpha,
                        const ShardOp& op, int64_t ro_s, int64_t co_s, const T* A_local, int64_t lda, int mode,
                        cudaStream_t st, T** partial_out) {
    RB_REQUIRE(c != nullptr);
    RB_REQUIRE(mode == 0 || mode == 1);
    RB_REQUIRE(layout == 'R' || layout == 'C');
    RB_REQUIRE(opS == 'N' || opS == 'T');
    RB_REQUIRE(d >= 0 && n >= 0 && m_total >= 0);
    if (mode == 0) RB_REQUIRE((d * n) % c->nranks == 0);       // reduce-scatter hands every rank d*n/nranks entries
    int64_t start, count;
    mshard_block(m_total, c->rank, c->nranks, &start, &count);
    T* W = (T*) comm_partial(c, (size_t) (d * n > 0 ? d * n : 1) * sizeof(T), st);
    if (!W) return fail_cuda(cudaErrorMemoryAllocation, "partial-product buffer of the communicator");
    *partial_out = W;
    if (d * n == 0) return 0;
    if (count == 0) { RB_CUDA(cudaMemsetAsync(W, 0, (size_t) (d * n) * sizeof(T), st)); return 0; }
    // the contraction index of op(S) runs along the columns of S for opS = N, along its rows for opS = T
    const int64_t ro = ro_s + (opS == 'T' ? start : 0), co = co_s + (opS == 'N' ? start : 0);
    const int64_t ldw = (layout == 'C') ? d : n;
    if (op.vec_nnz > 0)
        return Skge<T>::left_saso(layout, opS, opA, d, n, count, alpha, op.D_rows, op.D_cols, op.vec_nnz, op.ctr, op.key, ro, co,
                                  A_local, lda, (T) 0, W, ldw, (void*) st);
    return Skge<T>::left(layout, opS, opA, d, n, count, alpha, op.D_rows, op.D_cols, op.family, op.axis, op.ctr, op.key, ro, co,
                         A_local, lda, (T) 0, W, ldw, (void*) st);
}

// Phase 2: the exchange step, then beta * B_out.
template <typename T>
static int mshard_exchange(NcclApi* nc, rb_comm* c, int64_t d, int64_t n, T beta, T* W, T* B_out, int mode, cudaStream_t st) {
    const int64_t total = d * n;
    if (total == 0) return 0;
    const int64_t cnt = (mode == 0) ? total / c->nranks : total;
    T* slice = (mode == 0) ? W + (int64_t) c->rank * cnt : W;
    if (c->nranks == 1) {
        if (beta == (T) 0) RB_CUDA(cudaMemcpyAsync(B_out, W, (size_t) cnt * sizeof(T), cudaMemcpyDeviceToDevice, st));
    } else if (mode == 0) {
        RB_NCCL(nc, nc->ReduceScatter(W, beta == (T) 0 ? B_out : slice, (size_t) cnt, Skge<T>::nt, ncclSum, c->nccl, st));
    } else {
        RB_NCCL(nc, nc->AllReduce(W, beta == (T) 0 ? B_out : W, (size_t) cnt, Skge<T>::nt, ncclSum, c->nccl, st));
    }
    return 0;
}
template <typename T>
static int mshard_finish(rb_comm* c, int64_t d, int64_t n, T beta, T* W, T* B_out, int mode, cudaStream_t st) {
    const int64_t total = d * n;
    if (total == 0 || beta == (T) 0) return 0;
    const int64_t cnt = (mode == 0) ? total / c->nranks : total;
    T* slice = (mode == 0) ? W + (int64_t) c->rank * cnt : W;
    int64_t g = (cnt + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    axpby_kernel<T><<<(unsigned) g, 256, 0, st>>>(cnt, beta, B_out, slice);
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int lskge3_mshard(rb_comm* c, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m_total, T alpha,
                         const ShardOp& op, int64_t ro_s, int64_t co_s, const T* A_local, int64_t lda, T beta, T* B_out,
                         int mode, void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    RB_REQUIRE(c != nullptr);
    NcclApi* nc = nullptr;
    if (c->nranks > 1) {
        nc = nccl_api();
        if (!nc) return fail_nccl(nullptr, ncclSystemError, "dlopen(libnccl.so.2)");
    }
    RB_REQUIRE(B_out != nullptr || d * n == 0);
    int dev = 0;
    RB_CUDA(cudaGetDevice(&dev));
    RB_REQUIRE(dev == c->device);                 // the caller selects the communicator's device (cudaSetDevice)
    T* W = nullptr;
    int rc = mshard_local<T>(c, layout, opS, opA, d, n, m_total, alpha, op, ro_s, co_s, A_local, lda, mode, st, &W);
    if (rc) return rc;
    rc = mshard_exchange<T>(nc, c, d, n, beta, W, B_out, mode, st);
    if (rc) return rc;
    return mshard_finish<T>(c, d, n, beta, W, B_out, mode, st);
}

// One host thread driving all GPUs of a single-process communicator set (rb_comm_init).
template <typename T>
static int lskge3_mshard_all(int ndev, rb_comm* const* comms, char layout, char opS, char opA, int64_t d, int64_t n,
                             int64_t m_total, T alpha, const ShardOp& op, int64_t ro_s, int64_t co_s, const T* const* A_local,
                             const int64_t* lda, T beta, T* const* B_out, int mode, void* const* streams) {
    RB_REQUIRE(ndev >= 1 && comms != nullptr && A_local != nullptr && lda != nullptr && B_out != nullptr);
    NcclApi* nc = nullptr;
    if (ndev > 1) {
        nc = nccl_api();
        if (!nc) return fail_nccl(nullptr, ncclSystemError, "dlopen(libnccl.so.2)");
    }
    int cur = 0;
    RB_CUDA(cudaGetDevice(&cur));
    std::vector<T*> W((size_t) ndev, nullptr);
    int rc = 0;
    for (int g = 0; g < ndev && !rc; ++g) {                 // partial products: asynchronous, one stream per GPU
        RB_REQUIRE(comms[g] != nullptr && comms[g]->nranks == ndev && comms[g]->rank == g);
        cudaSetDevice(comms[g]->device);
        rc = mshard_local<T>(comms[g], layout, opS, opA, d, n, m_total, alpha, op, ro_s, co_s, A_local[g], lda[g], mode,
                             streams ? (cudaStream_t) streams[g] : nullptr, &W[g]);
    }
    if (!rc && ndev > 1) {
        ncclResult_t r = nc->GroupStart();
        if (r != ncclSuccess) rc = fail_nccl(nc, r, "ncclGroupStart");
        for (int g = 0; g < ndev && !rc; ++g) {
            cudaSetDevice(comms[g]->device);
            rc = mshard_exchange<T>(nc, comms[g], d, n, beta, W[g], B_out[g], mode, streams ? (cudaStream_t) streams[g] : nullptr);
        }
        r = nc->GroupEnd();
        if (!rc && r != ncclSuccess) rc = fail_nccl(nc, r, "ncclGroupEnd");
    } else if (!rc) {
        cudaSetDevice(comms[0]->device);
        rc = mshard_exchange<T>(nc, comms[0], d, n, beta, W[0], B_out[0], mode, streams ? (cudaStream_t) streams[0] : nullptr);
    }
    for (int g = 0; g < ndev && !rc; ++g) {
        cudaSetDevice(comms[g]->device);
        rc = mshard_finish<T>(comms[g], d, n, beta, W[g], B_out[g], mode, streams ? (cudaStream_t) streams[g] : nullptr);
    }
    cudaSetDevice(cur);
    return rc;
}

}  // namespace rb

using namespace rb;

extern "C" {

int rb_comm_unique_id(void* id128) {
    RB_REQUIRE(id128 != nullptr);
    static_assert(sizeof(ncclUniqueId) == RB_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    NcclApi* nc = nccl_api();
    if (!nc) return fail_nccl(nullptr, ncclSystemError, "dlopen(libnccl.so.2)");
    ncclUniqueId id;
    RB_NCCL(nc, nc->GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof id);
    return 0;
}

int rb_comm_init_rank(int nranks, int rank, const void* id128, rb_comm_t* comm) {
    RB_REQUIRE(comm != nullptr);
    RB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks);
    int dev = 0;
    RB_CUDA(cudaGetDevice(&dev));
    rb_comm* c = new rb_comm;
    c->nranks = nranks; c->rank = rank; c->device = dev;
    if (nranks > 1) {
        NcclApi* nc = nccl_api();
        if (!nc) { delete c; return fail_nccl(nullptr, ncclSystemError, "dlopen(libnccl.so.2)"); }
        if (!id128) { delete c; return fail("(id128 != nullptr) was required, but did not hold, in function rb_comm_init_rank"); }
        ncclUniqueId id;
        std::memcpy(&id, id128, sizeof id);
        ncclResult_t r = nc->CommInitRank(&c->nccl, nranks, id, rank);
        if (r != ncclSuccess) { delete c; return fail_nccl(nc, r, "ncclCommInitRank"); }
    }
    *comm = c;
    return 0;
}

int rb_comm_init(int ndev, const int* devices, rb_comm_t* comms) {
    RB_REQUIRE(ndev >= 1 && comms != nullptr);
    int have = 0;
    RB_CUDA(cudaGetDeviceCount(&have));
    std::vector<int> devs((size_t) ndev);
    for (int g = 0; g < ndev; ++g) {
        devs[g] = devices ? devices[g] : g;
        RB_REQUIRE(devs[g] >= 0 && devs[g] < have);
    }
    std::vector<ncclComm_t> nc_comms((size_t) ndev, nullptr);
    if (ndev > 1) {
        NcclApi* nc = nccl_api();
        if (!nc) return fail_nccl(nullptr, ncclSystemError, "dlopen(libnccl.so.2)");
        RB_NCCL(nc, nc->CommInitAll(nc_comms.data(), ndev, devs.data()));
    }
    for (int g = 0; g < ndev; ++g) {
        rb_comm* c = new rb_comm;
        c->nranks = ndev; c->rank = g; c->device = devs[g]; c->nccl = nc_comms[g];
        comms[g] = c;
    }
    return 0;
}

int rb_comm_destroy(rb_comm_t c) {
    if (!c) return 0;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(c->device);
    if (c->partial) { cudaDeviceSynchronize(); cudaFree(c->partial); }
    if (c->nccl) { NcclApi* nc = nccl_api(); if (nc) nc->CommDestroy(c->nccl); }
    cudaSetDevice(cur);
    cudaGetLastError();
    delete c;
    return 0;
}

int rb_comm_info(rb_comm_t c, int64_t info[4]) {
    RB_REQUIRE(c != nullptr && info != nullptr);
    info[0] = c->nranks; info[1] = c->rank; info[2] = c->device;
    info[3] = 0;
    if (NcclApi* nc = nccl_api()) { int v = 0; if (nc->GetVersion && nc->GetVersion(&v) == ncclSuccess) info[3] = v; }
    return 0;
}

int rb_mshard_block(int64_t m_total, int nranks, int rank, int64_t* start, int64_t* count) {
    RB_REQUIRE(m_total >= 0 && nranks >= 1 && rank >= 0 && rank < nranks && start != nullptr && count != nullptr);
    mshard_block(m_total, rank, nranks, start, count);
    return 0;
}

#define RB_DEF_MSHARD(T, sfx)                                                                                            \
    int rb_lskge3_mshard_##sfx(rb_comm_t comm, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m_total,   \
                               T alpha, int64_t D_rows, int64_t D_cols, char family, char major_axis,                    \
                               const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s, int64_t co_s,                 \
                               const T* A_local, int64_t lda, T beta, T* B_out, int mode, void* stream) {                \
        const ShardOp op{D_rows, D_cols, family, major_axis, 0, ctr, key};                                               \
        return lskge3_mshard<T>(comm, layout, opS, opA, d, n, m_total, alpha, op, ro_s, co_s, A_local, lda, beta, B_out, \
                                mode, stream);                                                                           \
    }                                                                                                                    \
    int rb_lskge3_mshard_all_##sfx(int ndev, const rb_comm_t* comms, char layout, char opS, char opA, int64_t d,         \
                                   int64_t n, int64_t m_total, T alpha, int64_t D_rows, int64_t D_cols, char family,     \
                                   char major_axis, const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s,          \
                                   int64_t co_s, const T* const* A_local, const int64_t* lda, T beta, T* const* B_out,   \
                                   int mode, void* const* streams) {                                                     \
        const ShardOp op{D_rows, D_cols, family, major_axis, 0, ctr, key};                                               \
        return lskge3_mshard_all<T>(ndev, comms, layout, opS, opA, d, n, m_total, alpha, op, ro_s, co_s, A_local, lda,   \
                                    beta, B_out, mode, streams);                                                         \
    }                                                                                                                    \
    int rb_lskges_mshard_##sfx(rb_comm_t comm, char layout, char opS, char opA, int64_t d, int64_t n, int64_t m_total,   \
                               T alpha, int64_t D_rows, int64_t D_cols, int64_t vec_nnz, const uint32_t ctr[4],          \
                               const uint32_t key[2], int64_t ro_s, int64_t co_s, const T* A_local, int64_t lda, T beta, \
                               T* B_out, int mode, void* stream) {                                                       \
        RB_REQUIRE(vec_nnz >= 1);                                                                                        \
        const ShardOp op{D_rows, D_cols, 'U', 'S', vec_nnz, ctr, key};                                                   \
        return lskge3_mshard<T>(comm, layout, opS, opA, d, n, m_total, alpha, op, ro_s, co_s, A_local, lda, beta, B_out, \
                                mode, stream);                                                                           \
    }                                                                                                                    \
    int rb_lskges_mshard_all_##sfx(int ndev, const rb_comm_t* comms, char layout, char opS, char opA, int64_t d,         \
                                   int64_t n, int64_t m_total, T alpha, int64_t D_rows, int64_t D_cols, int64_t vec_nnz, \
                                   const uint32_t ctr[4], const uint32_t key[2], int64_t ro_s, int64_t co_s,             \
                                   const T* const* A_local, const int64_t* lda, T beta, T* const* B_out, int mode,       \
                                   void* const* streams) {                                                               \
        RB_REQUIRE(vec_nnz >= 1);                                                                                        \
        const ShardOp op{D_rows, D_cols, 'U', 'S', vec_nnz, ctr, key};                                                   \
        return lskge3_mshard_all<T>(ndev, comms, layout, opS, opA, d, n, m_total, alpha, op, ro_s, co_s, A_local, lda,   \
                                    beta, B_out, mode, streams);                                                         \
    }
RB_DEF_MSHARD(float, f32)
RB_DEF_MSHARD(double, f64)

}  // extern "C"
