// Drop-in check of the header-only C++ layer (include/RandBLAS.hh) against librandblas_b200.so, written the way
// the reference's own gtest suites use the API (test/test_datastructures/test_denseskop.cc,
// test/test_matmul_cores/linop_common.hh:309-389, test_sparseskop.cc:64-117, test_sketch_vector.cc, test_r123.cc).
//   test_dropin --host : host-side checks only (no GPU needed): state arithmetic, distributions, error behaviour
//   test_dropin        : everything; needs a CUDA device. Buffers are plain host memory, as a caller of the CPU
//                        reference would pass them.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <set>
#include <vector>
#include <cuda_runtime_api.h>   // device buffers of the multi-GPU check only
#include "RandBLAS.hh"
#include "RandBLAS/sparse_data/random_matrix.hh"   // as in the reference: not part of the aggregate header

using namespace RandBLAS;
static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

template <typename F>
static bool throws_error(F f) {
    try { f(); } catch (const RandBLAS::Error&) { return true; } catch (...) { return false; }
    return false;
}

static void host_checks() {
    // RNGState(uint64) puts the integer into the key, little-endian limbs (test_r123.cc:679-698)
    RNGState<> s(0x100000002ull);
    CHECK(s.key[0] == 2u && s.key[1] == 1u);
    CHECK(s.counter[0] == 0u && s.counter[3] == 0u);
    // counter carries (test_r123.cc:714-797)
    RNGState<> t;
    t.counter.v[0] = 0xffffffffu; t.counter.v[1] = 0xffffffffu;
    t.counter.incr(1);
    CHECK(t.counter[0] == 0u && t.counter[1] == 0u && t.counter[2] == 1u);
    // DenseDist fields (dense_skops.hh:231-350)
    DenseDist D(7, 20);
    CHECK(D.dim_major == 20 && D.dim_minor == 7 && D.natural_layout == blas::Layout::RowMajor);
    CHECK(D.family == ScalarDist::Gaussian && D.major_axis == Axis::Long);
    CHECK(std::fabs(D.isometry_scale - 1.0 / std::sqrt(7.0)) < 1e-15);
    DenseDist Dt(20, 7, ScalarDist::Uniform, Axis::Short);
    CHECK(Dt.dim_major == 7 && Dt.natural_layout == blas::Layout::RowMajor);
    CHECK(throws_error([] { DenseDist bad(0, 5); }));
    // next_state of an operator = seed + dim_minor * ceil(dim_major / 4) (dense_skops.hh:172-185)
    DenseSkOp<float> S(D, RNGState<>(42));
    CHECK(S.next_state.counter[0] == 7u * 5u && S.next_state.key == S.seed_state.key);
    CHECK(S.buff == nullptr && S.own_memory && S.layout == blas::Layout::RowMajor);
    // SparseDist (sparse_skops.hh:131-246)
    SparseDist Ds(15, 7, 3);
    CHECK(Ds.major_axis == Axis::Short && Ds.dim_major == 7 && Ds.dim_minor == 15 && Ds.full_nnz == 45);
    CHECK(throws_error([] { SparseDist bad(4, 9, 5); }));       // vec_nnz > dim_major
    SparseSkOp<double, DefaultRNG, int32_t> Ss(Ds, RNGState<>(1));
    CHECK(Ss.nnz < 0 && Ss.next_state.counter[0] == 45u);
    // full-operator overloads check dimensions before anything else (skge.hh:1089-1095)
    std::vector<float> A(20 * 3), B(8 * 3);
    CHECK(throws_error([&] { sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, 8, 3, 20, 1.0f, S,
                                            A.data(), 20, 0.0f, B.data(), 8); }));
}

template <typename T>
static void device_checks() {
    const T tol = sizeof(T) == 4 ? T(1e-5) : T(1e-12);
    RNGState<> seed(1997);
    // --- fill_dense(S): allocation on demand, submatrix == slice of the full matrix (test_denseskop.cc:162-298)
    DenseDist D(11, 53, ScalarDist::Uniform);
    DenseSkOp<T> S(D, seed);
    fill_dense(S);
    CHECK(S.buff != nullptr);
    std::vector<T> sub(4 * 10);
    RNGState<> nxt = fill_dense_unpacked(blas::Layout::RowMajor, D, 4, 10, 3, 7, sub.data(), seed);
    bool same = true;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 10; ++j) same = same && sub[i * 10 + j] == S.buff[(3 + i) * 53 + 7 + j];
    CHECK(same);
    (void) nxt;
    T mx = 0;
    for (int64_t i = 0; i < 11 * 53; ++i) mx = std::max(mx, std::fabs(S.buff[i]));
    CHECK(mx <= T(1.7320509) && mx > T(1.5));          // uniform on [-sqrt 3, sqrt 3]
    // fill_dense(D, buff, seed) returns the operator's next_state (test_denseskop.cc:443-465)
    std::vector<T> full(11 * 53);
    RNGState<> n2 = fill_dense(D, full.data(), seed);
    CHECK(n2 == S.next_state);
    CHECK(std::memcmp(full.data(), S.buff, full.size() * sizeof(T)) == 0);

    // --- operator applied to the identity reproduces the operator (linop_common.hh:309-389), unfilled operator
    const int64_t d = 11, m = 53;
    DenseSkOp<T> S0(D, seed);
    std::vector<T> I(m * m, T(0)), B(d * m, T(0));
    for (int64_t i = 0; i < m; ++i) I[i * m + i] = T(1);
    sketch_general(blas::Layout::RowMajor, blas::Op::NoTrans, blas::Op::NoTrans, d, m, m, T(1), S0, I.data(), m, T(0),
                   B.data(), m);
    CHECK(S0.buff == nullptr);                         // an unfilled operator is not modified (skge.hh:174-181)
    double num = 0, den = 0;
    for (int64_t i = 0; i < d * m; ++i) { num += (B[i] - full[i]) * (B[i] - full[i]); den += full[i] * full[i]; }
    CHECK(std::sqrt(num / den) < tol);
    // right sketch with the transposed operator: I * S^T (skge.hh:947-968)
    std::vector<T> Bt(m * d, T(0));
    sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::Trans, m, d, m, T(1), I.data(), m, S0, T(0),
                   Bt.data(), m);
    num = 0;
    for (int64_t i = 0; i < d; ++i)
        for (int64_t j = 0; j < m; ++j) { double e = Bt[j + i * m] - full[i * m + j]; num += e * e; }
    CHECK(std::sqrt(num / den) < tol);

    // --- sketch_vector == row sums of S applied to ones (skve.hh:141-164)
    std::vector<T> x(m, T(1)), y(d, T(0));
    sketch_vector(blas::Op::NoTrans, T(1), S0, x.data(), 1, T(0), y.data(), 1);
    bool ok = true;
    for (int64_t i = 0; i < d; ++i) {
        double r = 0;
        for (int64_t j = 0; j < m; ++j) r += full[i * m + j];
        ok = ok && std::fabs(y[i] - r) <= 1e3 * tol * (1 + std::fabs(r));
    }
    CHECK(ok);

    // --- SASO: exactly vec_nnz distinct row indices per column, values +-1 (test_sparseskop.cc:64-117)
    SparseDist Ds(7, 200, 3);
    SparseSkOp<T, DefaultRNG, int32_t> Ss(Ds, seed);
    fill_sparse(Ss);
    CHECK(Ss.nnz == 600);
    ok = true;
    for (int64_t c = 0; c < 200; ++c) {
        std::set<int> rows;
        for (int j = 0; j < 3; ++j) {
            rows.insert(Ss.rows[3 * c + j]);
            ok = ok && Ss.cols[3 * c + j] == c && std::fabs(Ss.vals[3 * c + j]) == T(1);
            ok = ok && Ss.rows[3 * c + j] >= 0 && Ss.rows[3 * c + j] < 7;
        }
        ok = ok && rows.size() == 3;
    }
    CHECK(ok);
    // sampled (COO arrays) and unsampled (regenerated in the kernel) operators give the same product
    std::vector<T> A(200 * 5), B1(7 * 5, T(0)), B2(7 * 5, T(0));
    for (size_t i = 0; i < A.size(); ++i) A[i] = T((int) (i * 2654435761u % 1000)) / T(500) - T(1);
    SparseSkOp<T, DefaultRNG, int32_t> Su(Ds, seed);
    sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, 7, 5, 200, T(1), Ss, A.data(), 200, T(0),
                   B1.data(), 7);
    sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, 7, 5, 200, T(1), Su, A.data(), 200, T(0),
                   B2.data(), 7);
    CHECK(Su.nnz < 0);
    ok = true;
    for (int i = 0; i < 35; ++i) ok = ok && std::fabs(B1[i] - B2[i]) <= 100 * tol;
    CHECK(ok);

    // --- LASO (Axis::Long): every row of a wide operator holds at most vec_nnz entries, values +-sqrt(count),
    //     and sampled / unsampled operators give the same product (test_sparseskop.cc:119-146)
    SparseDist Dl(7, 200, 4, Axis::Long);
    SparseSkOp<T, DefaultRNG, int64_t> Sl(Dl, seed);
    fill_sparse(Sl);
    CHECK(Sl.nnz > 0 && Sl.nnz <= 28);
    ok = true;
    double sumsq = 0;
    for (int64_t e = 0; e < Sl.nnz; ++e) {
        ok = ok && Sl.rows[e] >= 0 && Sl.rows[e] < 7 && Sl.cols[e] >= 0 && Sl.cols[e] < 200;
        if (e > 0) ok = ok && Sl.rows[e] >= Sl.rows[e - 1];
        sumsq += (double) Sl.vals[e] * (double) Sl.vals[e];
    }
    CHECK(ok);
    CHECK(std::fabs(sumsq - 28.0) < 1e-4);                        // sqrt(count)^2 summed = draws = vec_nnz * dim_minor
    std::vector<T> B3(7 * 5, T(0)), B4(7 * 5, T(0));
    SparseSkOp<T, DefaultRNG, int64_t> Slu(Dl, seed);
    sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, 7, 5, 200, T(1), Sl, A.data(), 200, T(0),
                   B3.data(), 7);
    sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, 7, 5, 200, T(1), Slu, A.data(), 200, T(0),
                   B4.data(), 7);
    CHECK(Slu.nnz < 0);
    ok = true;
    for (int i = 0; i < 35; ++i) ok = ok && std::fabs(B3[i] - B4[i]) <= 100 * tol;
    CHECK(ok);

    // --- sketch_sparse applied to a sparse identity reproduces S (test_sketch_sparse.cc, COO by design)
    std::vector<T> ones(m, T(1));
    std::vector<int64_t> idx(m);
    for (int64_t i = 0; i < m; ++i) idx[i] = i;
    COOMatrix<T> Isp(m, m, m, ones.data(), idx.data(), idx.data());
    std::vector<T> Bs(d * m, T(0));
    sketch_sparse(blas::Layout::RowMajor, blas::Op::NoTrans, blas::Op::NoTrans, d, m, m, T(1), S0, 0, 0, Isp, T(0), Bs.data(),
                  m);
    num = 0;
    for (int64_t i = 0; i < d * m; ++i) num += (Bs[i] - full[i]) * (Bs[i] - full[i]);
    CHECK(std::sqrt(num / den) < tol);

    // --- sparse_data::left_spmm with the sparse identity in CSR form reproduces the dense operand (spmm_dispatch.hh)
    {
        std::vector<int64_t> rowptr(m + 1);
        for (int64_t i = 0; i <= m; ++i) rowptr[i] = i;
        CSRMatrix<T> Icsr(m, m, m, ones.data(), rowptr.data(), idx.data());
        std::vector<T> Cs(m * d, T(0));
        // C(m x d) = I(m x m) * full^T: op(B) = full(d x m)^T, RowMajor
        sparse_data::left_spmm(blas::Layout::RowMajor, blas::Op::NoTrans, blas::Op::Trans, m, d, m, T(1), Icsr, 0, 0,
                               full.data(), m, T(0), Cs.data(), d);
        ok = true;
        for (int64_t i = 0; i < d; ++i)
            for (int64_t j = 0; j < m; ++j) ok = ok && Cs[j * d + i] == full[i * m + j];
        CHECK(ok);
    }

    // --- format conversions (sparse_data/conversions.hh): COO -> CSR -> COO round trip of a small unsorted matrix
    {
        int64_t r[5] = {2, 0, 1, 2, 0}, c[5] = {1, 3, 0, 0, 1};
        T v[5] = {T(5), T(2), T(3), T(4), T(1)};
        COOMatrix<T> Acoo(3, 4, 5, v, r, c);
        CSRMatrix<T> Acsr(3, 4);
        sparse_data::coo_to_csr(Acoo, Acsr);
        const int64_t want_ptr[4] = {0, 2, 3, 5}, want_col[5] = {1, 3, 0, 0, 1};
        const T want_val[5] = {T(1), T(2), T(3), T(4), T(5)};
        ok = Acsr.nnz == 5;
        for (int i = 0; i < 4; ++i) ok = ok && Acsr.rowptr[i] == want_ptr[i];
        for (int i = 0; i < 5; ++i) ok = ok && Acsr.colidxs[i] == want_col[i] && Acsr.vals[i] == want_val[i];
        CHECK(ok);
        COOMatrix<T> Aback(3, 4);
        sparse_data::csr_to_coo(Acsr, Aback);
        const int64_t want_row[5] = {0, 0, 1, 2, 2};
        ok = Aback.nnz == 5;
        for (int i = 0; i < 5; ++i) ok = ok && Aback.rows[i] == want_row[i] && Aback.cols[i] == want_col[i];
        CHECK(ok);
    }

    // --- sketch_symmetric (sksy.hh:159-176): symmetric input sketches, asymmetric input throws
    {
        std::vector<T> Asym(m * m), Bsym(m * d, T(0));
        for (int64_t i = 0; i < m; ++i)
            for (int64_t j = 0; j < m; ++j) Asym[i * m + j] = T(1) / T(1 + i + j);
        DenseSkOp<T> Sr(DenseDist(m, d), seed);
        sketch_symmetric(blas::Layout::RowMajor, m, d, T(1), Asym.data(), m, Sr, 0, 0, T(0), Bsym.data(), d);
        Asym[3] += T(0.5);
        CHECK(throws_error([&] { sketch_symmetric(blas::Layout::RowMajor, m, d, T(1), Asym.data(), m, Sr, 0, 0, T(0),
                                                  Bsym.data(), d); }));
    }

    // --- index sampling (util.hh:459-560), written like test/test_basic_rng/test_discrete.cc:72-212
    {
        const int64_t N = 100, ns = N * N;
        std::vector<int64_t> samples(ns, -1);
        RNGState<> st(0);
        sample_indices_iid_uniform(N, ns, samples.data(), st);
        ok = true;
        for (auto v : samples) ok = ok && v >= 0 && v < N;
        CHECK(ok);                                                  // test_iid_uniform_smoke
        // degenerate distributions (test_discrete.cc:136-167)
        std::vector<float> cdf(N, 0.0f);
        for (int i = 0; i < N; i += 2) cdf[i] = 1.0f / ((float) i + 1.0f);
        cdf[10] = 0.0f;
        weights_to_cdf(N, cdf.data());
        CHECK(cdf[N - 1] == 1.0f || cdf[N - 2] == 1.0f);
        sample_indices_iid(N, cdf.data(), ns, samples.data(), st);
        ok = true;
        for (auto v : samples) ok = ok && !(v == 10 || v % 2 == 1);
        CHECK(ok);
        std::fill(cdf.begin(), cdf.end(), 0.0f);
        cdf[17] = 99.0f;
        cdf[3] = -std::numeric_limits<float>::epsilon() / 10;       // clipped without error
        weights_to_cdf(N, cdf.data());
        sample_indices_iid(N, cdf.data(), ns, samples.data(), st);
        ok = true;
        for (auto v : samples) ok = ok && v == 17;
        CHECK(ok);
        cdf[5] = -1.0f;
        CHECK(throws_error([&] { weights_to_cdf(N, cdf.data()); }));   // a weight below error_if_below (util.hh:465)
        // state updates (test_discrete.cc:169-212)
        RNGState<> sd;
        sd.counter.incr(3456);
        std::vector<int> un(34);
        auto s1 = sample_indices_iid_uniform(40, 17, un.data(), sd);
        auto s2 = sample_indices_iid_uniform(40, 17, un.data(), s1);
        auto t1 = sample_indices_iid_uniform(40, 34, un.data(), sd);
        CHECK(s2.counter.v[0] - 3456 == 2 * (s1.counter.v[0] - 3456));
        CHECK(t1.counter.v[0] <= s2.counter.v[0] && s2.counter.v[0] <= t1.counter.v[0] + 1);
        std::vector<float> ucdf(29, 1.0f);
        weights_to_cdf(29, ucdf.data());
        auto c1 = sample_indices_iid(29, ucdf.data(), 13, un.data(), sd);
        auto c2 = sample_indices_iid(29, ucdf.data(), 13, un.data(), c1);
        auto ct = sample_indices_iid(29, ucdf.data(), 26, un.data(), sd);
        CHECK(c2.counter.v[0] - 3456 == 2 * (c1.counter.v[0] - 3456));
        CHECK(ct.counter.v[0] <= c2.counter.v[0] && c2.counter.v[0] <= ct.counter.v[0] + 1);
        // Rademacher form: signs are +-1, two words per sample
        std::vector<T> rad(34);
        auto r1 = sample_indices_iid_uniform<T, int, true>(40, 34, un.data(), rad.data(), sd);
        ok = r1.counter.v[0] == 3456u + 17u;
        for (auto v : rad) ok = ok && (v == T(1) || v == T(-1));
        CHECK(ok);
    }

    // --- argument errors surface as RandBLAS::Error before data is touched (skge.hh:183-192)
    CHECK(throws_error([&] { sketch_general(blas::Layout::RowMajor, blas::Op::NoTrans, blas::Op::NoTrans, d, m, m, T(1), S0, 1,
                                            0, I.data(), m, T(0), B.data(), m); }));
    CHECK(throws_error([&] { sketch_general(blas::Layout::RowMajor, blas::Op::NoTrans, blas::Op::NoTrans, d, m, m, T(1), S0, 0,
                                            0, I.data(), m - 1, T(0), B.data(), m); }));
}

// --- multi-GPU left sketch through the drop-in header (include/RandBLAS/multi_gpu.hh): all GPUs of this process
// (one is enough: a communicator set of one rank needs no NCCL), device buffers, against the single-call sketch.
template <typename T>
static void multi_gpu_checks() {
    using namespace RandBLAS::multi_gpu;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) { CHECK(false); return; }
    const int nd = have >= 2 ? 2 : 1;
    const T tol = sizeof(T) == 4 ? T(1e-5) : T(1e-12);
    const int64_t d = 64, n = 24, m = 4003;
    RNGState<> seed(1997);
    DenseSkOp<T> S(DenseDist(d, m), seed);
    std::vector<T> A(m * n);                                    // RowMajor m x n
    for (int64_t i = 0; i < m * n; ++i) A[i] = T(((i * 2654435761u) % 2001) / 1000.0 - 1.0);
    std::vector<T> Bfull(d * n, T(0));
    sketch_general(blas::Layout::RowMajor, blas::Op::NoTrans, blas::Op::NoTrans, d, n, m, T(1), S, 0, 0, A.data(), n, T(0),
                   Bfull.data(), n);
    Communicators comms(nd);
    CHECK(comms.size() == nd);
    std::vector<T*> dA(nd, nullptr), dB(nd, nullptr);
    std::vector<const T*> cA(nd, nullptr);
    std::vector<int64_t> lda(nd, n);
    int64_t covered = 0;
    for (int g = 0; g < nd; ++g) {
        auto [start, count] = mshard_block(m, nd, g);
        CHECK(start == covered && (start % 4 == 0 || start == m));
        covered += count;
        cudaSetDevice(comms.device(g));
        cudaMalloc((void**) &dA[g], sizeof(T) * (size_t) (count > 0 ? count * n : 1));
        cudaMalloc((void**) &dB[g], sizeof(T) * (size_t) (d * n));
        cudaMemcpy(dA[g], A.data() + start * n, sizeof(T) * (size_t) (count * n), cudaMemcpyHostToDevice);
        cA[g] = dA[g];
    }
    CHECK(covered == m);
    for (Reduce mode : {Reduce::Scatter, Reduce::All}) {
        sketch_general_mshard(comms, blas::Layout::RowMajor, blas::Op::NoTrans, blas::Op::NoTrans, d, n, m, T(1), S, 0, 0, cA.data(),
                              lda.data(), T(0), dB.data(), mode);
        for (int g = 0; g < nd; ++g) {
            cudaSetDevice(comms.device(g));
            cudaDeviceSynchronize();
            const int64_t cnt = mode == Reduce::Scatter ? d * n / nd : d * n, off = mode == Reduce::Scatter ? g * cnt : 0;
            std::vector<T> got((size_t) cnt);
            cudaMemcpy(got.data(), dB[g], sizeof(T) * (size_t) cnt, cudaMemcpyDeviceToHost);
            double num = 0, den = 0;
            for (int64_t i = 0; i < cnt; ++i) { double e = (double) got[i] - (double) Bfull[off + i]; num += e * e; den += (double) Bfull[off + i] * Bfull[off + i]; }
            CHECK(std::sqrt(num / den) < tol);
        }
    }
    // the same data against an unsampled SASO operator (sparse::lskges per GPU, same collective)
    {
        SparseSkOp<T> Ssp(SparseDist(d, m, 4), seed);
        std::vector<T> Bsp(d * n, T(0));
        sketch_general(blas::Layout::RowMajor, blas::Op::NoTrans, blas::Op::NoTrans, d, n, m, T(1), Ssp, 0, 0, A.data(), n, T(0),
                       Bsp.data(), n);
        sketch_general_mshard(comms, blas::Layout::RowMajor, blas::Op::NoTrans, blas::Op::NoTrans, d, n, m, T(1), Ssp, 0, 0,
                              cA.data(), lda.data(), T(0), dB.data(), Reduce::All);
        for (int g = 0; g < nd; ++g) {
            cudaSetDevice(comms.device(g));
            cudaDeviceSynchronize();
            std::vector<T> got((size_t) (d * n));
            cudaMemcpy(got.data(), dB[g], sizeof(T) * (size_t) (d * n), cudaMemcpyDeviceToHost);
            double num = 0, den = 0;
            for (int64_t i = 0; i < d * n; ++i) { double e = (double) got[i] - (double) Bsp[i]; num += e * e; den += (double) Bsp[i] * Bsp[i]; }
            CHECK(std::sqrt(num / den) < tol);
        }
        CHECK(Ssp.nnz < 0);                                     // the operator was not sampled by either call
    }
    for (int g = 0; g < nd; ++g) { cudaSetDevice(comms.device(g)); cudaFree(dA[g]); cudaFree(dB[g]); }
    cudaSetDevice(0);
    // a filled operator is refused, as is a communicator array of the wrong size
    std::printf("multi_gpu_checks<%s>: %d GPU(s)\n", sizeof(T) == 4 ? "float" : "double", nd);
}

// --- random sparse matrices (RandBLAS/sparse_data/random_matrix.hh:136-355), written like the reference's
// test_datastructures/test_spmats/test_random_matrix.cc: shapes, density, sortedness, determinism, state chaining
static void random_matrix_checks() {
    using namespace RandBLAS::sparse_data;
    const int64_t m = 300, n = 500;
    const double dens = 0.02;
    RNGState<> st(42);
    auto [A, next] = random_coo<double>(m, n, dens, st);
    CHECK(A.n_rows == m && A.n_cols == n && A.nnz > 0 && A.sort == NonzeroSort::CSR);
    CHECK(std::fabs((double) A.nnz - m * n * dens) < 6.0 * std::sqrt(m * n * dens));
    bool ok = true;
    for (int64_t e = 0; e < A.nnz; ++e) {
        ok = ok && A.rows[e] >= 0 && A.rows[e] < m && A.cols[e] >= 0 && A.cols[e] < n;
        if (e > 0) ok = ok && (A.rows[e] * n + A.cols[e] > A.rows[e - 1] * n + A.cols[e - 1]);     // strictly row-major sorted
    }
    CHECK(ok);
    CHECK(next.counter.v[0] == (uint32_t) (A.nnz / 2 + 1) && next.key == st.key);      // the stream stops in block nnz / 2
    auto [A2, next2] = random_coo<double>(m, n, dens, st);                               // same state, same matrix
    ok = A2.nnz == A.nnz && next2 == next;
    for (int64_t e = 0; e < A.nnz && ok; ++e) ok = A.rows[e] == A2.rows[e] && A.cols[e] == A2.cols[e] && A.vals[e] == A2.vals[e];
    CHECK(ok);
    auto [Af, nextf] = random_coo<float, int>(m, n, dens, st);                           // float values are the double ones rounded
    ok = Af.nnz == A.nnz && nextf == next;
    for (int64_t e = 0; e < A.nnz && ok; ++e) ok = Af.rows[e] == (int) A.rows[e] && Af.vals[e] == (float) A.vals[e];
    CHECK(ok);
    auto [R, nr] = random_csr<double>(m, n, dens, st);
    ok = R.nnz == A.nnz && R.rowptr[0] == 0 && R.rowptr[m] == A.nnz && nr == next;
    for (int64_t i = 0; i < m && ok; ++i)
        for (int64_t e = R.rowptr[i]; e < R.rowptr[i + 1] && ok; ++e) ok = A.rows[e] == i && R.colidxs[e] == A.cols[e];
    CHECK(ok);
    auto [C, nc] = random_csc<double>(m, n, dens, st);
    ok = C.n_rows == m && C.n_cols == n && C.colptr[0] == 0 && C.colptr[n] == C.nnz;
    for (int64_t j = 0; j < n && ok; ++j)
        for (int64_t e = C.colptr[j]; e + 1 < C.colptr[j + 1] && ok; ++e) ok = C.rowidxs[e] < C.rowidxs[e + 1];
    CHECK(ok);
    auto [E, ne] = random_coo<double>(m, n, 0.0, st);                                    // density 0: nothing drawn
    CHECK(E.nnz == 0 && ne == st);
    CHECK(throws_error([&] { random_coo<double>(m, n, 1.5, st); }));
    // a sketch of the random matrix through sketch_sparse, as the reference's examples do
    const int64_t d = 16;
    DenseSkOp<double> S(DenseDist(d, m), RNGState<>(7));
    std::vector<double> B(d * n, 0.0);
    sketch_sparse(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, d, n, m, 1.0, S, 0, 0, R, 0.0, B.data(), d);
    double nrm = 0;
    for (auto v : B) nrm += v * v;
    CHECK(nrm > 0);
}

int main(int argc, char** argv) {
    host_checks();
    if (!(argc > 1 && std::strcmp(argv[1], "--host") == 0)) {
        device_checks<float>();
        device_checks<double>();
        random_matrix_checks();
        multi_gpu_checks<float>();
        multi_gpu_checks<double>();
    }
    std::printf(failures ? "test_dropin: %d FAILURES\n" : "test_dropin: all checks passed\n", failures);
    return failures ? 1 : 0;
}
