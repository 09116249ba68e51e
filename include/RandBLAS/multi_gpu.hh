// randblas_b200 -- header-only drop-in layer, part 7: the multi-GPU form of the left sketch.
//
// The reference has no multi-device code; what it has is the blocked form of sketch_general that this distributes:
// row blocks of A against column blocks of S selected with (ro_s, co_s), block products accumulated into one B
// (RandBLAS/skge.hh:174-181; rtd/source/tutorial/sketch_updates.rst:198-213). Here block g lives on GPU g, every GPU
// regenerates only its columns of the operator from the shared seed state, and the d x n partial products are summed
// by one NCCL reduce-scatter (or all-reduce) inside librandblas_b200.so (rb_comm_* / rb_lskge3_mshard_*).
// A_local / B_out are DEVICE pointers; everything else keeps the meaning it has in sketch_general.
#pragma once
#include <utility>
#include <vector>
#include "dense_skops.hh"
#include "sparse_skops.hh"

namespace RandBLAS::multi_gpu {

enum class Reduce : int { Scatter = 0, All = 1 };

// rows [start, start + count) of op(A) owned by `rank`; starts are multiples of 4 (no Philox block is split)
inline std::pair<int64_t, int64_t> mshard_block(int64_t m_total, int nranks, int rank) {
    int64_t s = 0, c = 0;
    internal::check(rb_mshard_block(m_total, nranks, rank, &s, &c), __func__);
    return {s, c};
}

// All GPUs of one process (the reference's callers are single-process programs).
class Communicators {
   public:
    explicit Communicators(int ndev, const int* devices = nullptr) : comms_((size_t) ndev, nullptr) {
        randblas_require(ndev >= 1);
        internal::check(rb_comm_init(ndev, devices, comms_.data()), __func__);
    }
    ~Communicators() { for (auto c : comms_) rb_comm_destroy(c); }
    Communicators(const Communicators&) = delete;
    Communicators& operator=(const Communicators&) = delete;
    int size() const { return (int) comms_.size(); }
    const rb_comm_t* data() const { return comms_.data(); }
    int device(int g) const { int64_t info[4]; internal::check(rb_comm_info(comms_[(size_t) g], info), __func__); return (int) info[2]; }

   private:
    std::vector<rb_comm_t> comms_;
};

// One GPU of a one-process-per-GPU job: rank 0 creates the id, every rank constructs with it (its GPU current).
class RankCommunicator {
   public:
    struct UniqueId { char bytes[RB_COMM_ID_BYTES]; };
    static UniqueId unique_id() { UniqueId id; internal::check(rb_comm_unique_id(id.bytes), __func__); return id; }
    RankCommunicator(int nranks, int rank, const UniqueId* id) : comm_(nullptr) {
        internal::check(rb_comm_init_rank(nranks, rank, id ? id->bytes : nullptr, &comm_), __func__);
    }
    ~RankCommunicator() { rb_comm_destroy(comm_); }
    RankCommunicator(const RankCommunicator&) = delete;
    RankCommunicator& operator=(const RankCommunicator&) = delete;
    rb_comm_t get() const { return comm_; }

   private:
    rb_comm_t comm_;
};

namespace internal_mg {
inline int call(rb_comm_t c, char l, char oS, char oA, int64_t d, int64_t n, int64_t m, float alpha, int64_t Dr, int64_t Dc,
                char fam, char ax, const uint32_t* ctr, const uint32_t* key, int64_t ro, int64_t co, const float* A, int64_t lda,
                float beta, float* B, int mode, void* st) {
    return rb_lskge3_mshard_f32(c, l, oS, oA, d, n, m, alpha, Dr, Dc, fam, ax, ctr, key, ro, co, A, lda, beta, B, mode, st);
}
inline int call(rb_comm_t c, char l, char oS, char oA, int64_t d, int64_t n, int64_t m, double alpha, int64_t Dr, int64_t Dc,
                char fam, char ax, const uint32_t* ctr, const uint32_t* key, int64_t ro, int64_t co, const double* A, int64_t lda,
                double beta, double* B, int mode, void* st) {
    return rb_lskge3_mshard_f64(c, l, oS, oA, d, n, m, alpha, Dr, Dc, fam, ax, ctr, key, ro, co, A, lda, beta, B, mode, st);
}
inline int call_all(int nd, const rb_comm_t* c, char l, char oS, char oA, int64_t d, int64_t n, int64_t m, float alpha,
                    int64_t Dr, int64_t Dc, char fam, char ax, const uint32_t* ctr, const uint32_t* key, int64_t ro, int64_t co,
                    const float* const* A, const int64_t* lda, float beta, float* const* B, int mode, void* const* st) {
    return rb_lskge3_mshard_all_f32(nd, c, l, oS, oA, d, n, m, alpha, Dr, Dc, fam, ax, ctr, key, ro, co, A, lda, beta, B, mode, st);
}
inline int call_all(int nd, const rb_comm_t* c, char l, char oS, char oA, int64_t d, int64_t n, int64_t m, double alpha,
                    int64_t Dr, int64_t Dc, char fam, char ax, const uint32_t* ctr, const uint32_t* key, int64_t ro, int64_t co,
                    const double* const* A, const int64_t* lda, double beta, double* const* B, int mode, void* const* st) {
    return rb_lskge3_mshard_all_f64(nd, c, l, oS, oA, d, n, m, alpha, Dr, Dc, fam, ax, ctr, key, ro, co, A, lda, beta, B, mode, st);
}
inline int call_saso(rb_comm_t c, char l, char oS, char oA, int64_t d, int64_t n, int64_t m, float alpha, int64_t Dr, int64_t Dc,
                     int64_t k, const uint32_t* ctr, const uint32_t* key, int64_t ro, int64_t co, const float* A, int64_t lda,
                     float beta, float* B, int mode, void* st) {
    return rb_lskges_mshard_f32(c, l, oS, oA, d, n, m, alpha, Dr, Dc, k, ctr, key, ro, co, A, lda, beta, B, mode, st);
}
inline int call_saso(rb_comm_t c, char l, char oS, char oA, int64_t d, int64_t n, int64_t m, double alpha, int64_t Dr, int64_t Dc,
                     int64_t k, const uint32_t* ctr, const uint32_t* key, int64_t ro, int64_t co, const double* A, int64_t lda,
                     double beta, double* B, int mode, void* st) {
    return rb_lskges_mshard_f64(c, l, oS, oA, d, n, m, alpha, Dr, Dc, k, ctr, key, ro, co, A, lda, beta, B, mode, st);
}
inline int call_saso_all(int nd, const rb_comm_t* c, char l, char oS, char oA, int64_t d, int64_t n, int64_t m, float alpha,
                         int64_t Dr, int64_t Dc, int64_t k, const uint32_t* ctr, const uint32_t* key, int64_t ro, int64_t co,
                         const float* const* A, const int64_t* lda, float beta, float* const* B, int mode, void* const* st) {
    return rb_lskges_mshard_all_f32(nd, c, l, oS, oA, d, n, m, alpha, Dr, Dc, k, ctr, key, ro, co, A, lda, beta, B, mode, st);
}
inline int call_saso_all(int nd, const rb_comm_t* c, char l, char oS, char oA, int64_t d, int64_t n, int64_t m, double alpha,
                         int64_t Dr, int64_t Dc, int64_t k, const uint32_t* ctr, const uint32_t* key, int64_t ro, int64_t co,
                         const double* const* A, const int64_t* lda, double beta, double* const* B, int mode, void* const* st) {
    return rb_lskges_mshard_all_f64(nd, c, l, oS, oA, d, n, m, alpha, Dr, Dc, k, ctr, key, ro, co, A, lda, beta, B, mode, st);
}
}  // namespace internal_mg

// B = alpha * op(S[ro_s:, co_s:]) * op(A) + beta * B with the m rows of op(A) sharded over the GPUs of `comms`:
// A_local[g] (device memory of GPU g) holds rows mshard_block(m, size, g) with leading dimension lda[g]; B_out[g]
// receives GPU g's slice (Reduce::Scatter: d*n/size entries of the packed result) or the whole result (Reduce::All).
// Asynchronous: ordered on streams[g] (or each GPU's default stream); synchronise the devices to read B_out.
template <typename T, typename RNG>
inline void sketch_general_mshard(const Communicators& comms, blas::Layout layout, blas::Op opS, blas::Op opA, int64_t d, int64_t n,
                                  int64_t m, T alpha, const DenseSkOp<T, RNG>& S, int64_t ro_s, int64_t co_s,
                                  const T* const* A_local, const int64_t* lda, T beta, T* const* B_out, Reduce mode = Reduce::Scatter,
                                  void* const* streams = nullptr) {
    randblas_require(S.buff == nullptr);      // the operator is regenerated per GPU; a filled S is a single-GPU object
    internal::check(internal_mg::call_all(comms.size(), comms.data(), internal::to_char(layout), internal::to_char(opS),
                                          internal::to_char(opA), d, n, m, alpha, S.dist.n_rows, S.dist.n_cols, (char) S.dist.family,
                                          (char) S.dist.major_axis, S.seed_state.counter.v, S.seed_state.key.v, ro_s, co_s, A_local,
                                          lda, beta, B_out, (int) mode, streams),
                    __func__);
}

// The same call for ONE rank of a one-process-per-GPU job.
template <typename T, typename RNG>
inline void sketch_general_mshard(const RankCommunicator& comm, blas::Layout layout, blas::Op opS, blas::Op opA, int64_t d, int64_t n,
                                  int64_t m, T alpha, const DenseSkOp<T, RNG>& S, int64_t ro_s, int64_t co_s, const T* A_local,
                                  int64_t lda, T beta, T* B_out, Reduce mode = Reduce::Scatter, void* stream = nullptr) {
    randblas_require(S.buff == nullptr);
    internal::check(internal_mg::call(comm.get(), internal::to_char(layout), internal::to_char(opS), internal::to_char(opA), d, n, m,
                                      alpha, S.dist.n_rows, S.dist.n_cols, (char) S.dist.family, (char) S.dist.major_axis,
                                      S.seed_state.counter.v, S.seed_state.key.v, ro_s, co_s, A_local, lda, beta, B_out, (int) mode,
                                      stream),
                    __func__);
}

// The two calls above for an unsampled SASO operator (SparseDist with Axis::Short): sparse::lskges (RandBLAS/skge.hh:465-492) on
// every GPU's row block, the d x n partials summed by the same collective.
template <typename T, typename RNG, typename sint_t>
inline void sketch_general_mshard(const Communicators& comms, blas::Layout layout, blas::Op opS, blas::Op opA, int64_t d, int64_t n,
                                  int64_t m, T alpha, const SparseSkOp<T, RNG, sint_t>& S, int64_t ro_s, int64_t co_s,
                                  const T* const* A_local, const int64_t* lda, T beta, T* const* B_out, Reduce mode = Reduce::Scatter,
                                  void* const* streams = nullptr) {
    randblas_require(S.nnz < 0);              // the operator is regenerated per GPU; a sampled S is a single-GPU object
    randblas_require(S.dist.major_axis == Axis::Short);
    internal::check(internal_mg::call_saso_all(comms.size(), comms.data(), internal::to_char(layout), internal::to_char(opS),
                                               internal::to_char(opA), d, n, m, alpha, S.dist.n_rows, S.dist.n_cols, S.dist.vec_nnz,
                                               S.seed_state.counter.v, S.seed_state.key.v, ro_s, co_s, A_local, lda, beta, B_out,
                                               (int) mode, streams),
                    __func__);
}
template <typename T, typename RNG, typename sint_t>
inline void sketch_general_mshard(const RankCommunicator& comm, blas::Layout layout, blas::Op opS, blas::Op opA, int64_t d, int64_t n,
                                  int64_t m, T alpha, const SparseSkOp<T, RNG, sint_t>& S, int64_t ro_s, int64_t co_s,
                                  const T* A_local, int64_t lda, T beta, T* B_out, Reduce mode = Reduce::Scatter,
                                  void* stream = nullptr) {
    randblas_require(S.nnz < 0);
    randblas_require(S.dist.major_axis == Axis::Short);
    internal::check(internal_mg::call_saso(comm.get(), internal::to_char(layout), internal::to_char(opS), internal::to_char(opA), d, n,
                                           m, alpha, S.dist.n_rows, S.dist.n_cols, S.dist.vec_nnz, S.seed_state.counter.v,
                                           S.seed_state.key.v, ro_s, co_s, A_local, lda, beta, B_out, (int) mode, stream),
                    __func__);
}

}  // namespace RandBLAS::multi_gpu
