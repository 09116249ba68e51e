// Generic fused generate-and-multiply kernel (SIMT) for the dense-operator sketches.
//
// Replaces dense::lskge3 / rskge3 (RandBLAS/skge.hh:154-202, 307-355) for every combination of layout,
// opS, opA, offsets, alpha/beta and scalar type, including the GEMV-shaped sketch_vector
// (RandBLAS/skve.hh:141-164). The operator tile is regenerated from (key, counter, ro_s, co_s) into shared
// memory; S never touches HBM. When S.buff was filled by the caller (S_buff != null) the tile is read from
// it instead (the reference's blas::gemm path, skge.hh:194-200).
//
// This is the any-shape kernel: the tensor-core kernels (skge3_f32_tc.cu, skge3_f64_dmma.cu) take the
// large aligned cases and fall back here otherwise.
#include "common.cuh"
#include "kernels.h"

namespace rb {

namespace {

constexpr int TI = 64, TJ = 64, TK = 16, NT = 256;

template <typename T, bool GAUSS, bool FROM_MEM>
__global__ void __launch_bounds__(NT) dense_generic_kernel(const DenseProblem<T> p) {
    __shared__ __align__(16) double2 logtab[(GAUSS && !FROM_MEM) ? LOGF_TABLE_ENTRIES : 1];
    __shared__ T Xs[TK][TI + 4];
    __shared__ T Ys[TK][TJ + 4];
    if constexpr (GAUSS && !FROM_MEM) load_logf_table(logtab, p.gen.logtab);

    const int tid = threadIdx.x;
    const int64_t i0 = (int64_t) blockIdx.y * TI, j0 = (int64_t) blockIdx.x * TJ;
    const int ti = tid / 16, tj = tid % 16;   // 16 x 16 threads, 4 x 4 outputs each
    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = (T) 0;

    const bool u_along_k = p.uk == 1;
    for (int64_t k0 = 0; k0 < p.K; k0 += TK) {
        __syncthreads();
        // ---- operator tile X[i0:i0+TI, k0:k0+TK] -> Xs[k][i]
        if constexpr (FROM_MEM) {
            for (int e = tid; e < TI * TK; e += NT) {
                int ii, kk;
                if (u_along_k) { kk = e % TK; ii = e / TK; } else { ii = e % TI; kk = e / TI; }
                T x = (T) 0;
                if (i0 + ii < p.P && k0 + kk < p.K) {
                    int64_t v = p.v0 + (i0 + ii) * p.vi + (k0 + kk) * p.vk;
                    int64_t u = p.u0 + (i0 + ii) * p.ui + (k0 + kk) * p.uk;
                    x = p.S_buff[v * p.S_ld + u];
                }
                Xs[kk][ii] = x;
            }
        } else {
            // window of the tile in natural coordinates
            const int nvt = u_along_k ? TI : TK, nut = u_along_k ? TK : TI;
            const int64_t va = p.v0 + (u_along_k ? i0 : k0);
            const int64_t ua = p.u0 + (u_along_k ? k0 : i0);
            const int64_t blk_a = ua >> 2;
            const int nbt = (int) (((ua + nut - 1) >> 2) - blk_a + 1);
            const int64_t vlim = u_along_k ? (p.P - i0) : (p.K - k0);   // valid extent along v
            const int64_t ulim = u_along_k ? (p.K - k0) : (p.P - i0);   // valid extent along u
            for (int e = tid; e < nvt * nbt; e += NT) {
                const int vl = e / nbt, bl = e - vl * nbt;
                float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                if (vl < vlim) {
                    const Ctr128 c = ctr_add(p.gen.ctr, (uint64_t) ((va + vl) * p.gen.R + blk_a + bl));
                    f = transform4<GAUSS>(philox4x32_10(c, p.gen.key), logtab);
                }
                const float fl[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                for (int lane = 0; lane < 4; ++lane) {
                    const int64_t ul = ((blk_a + bl) << 2) + lane - ua;
                    if (ul < 0 || ul >= nut) continue;
                    T x = (vl < vlim && ul < ulim) ? finish_sample<T, GAUSS>(fl[lane]) : (T) 0;
                    if (u_along_k) Xs[ul][vl] = x; else Xs[vl][ul] = x;
                }
            }
        }
        // ---- data tile Y[k0:k0+TK, j0:j0+TJ] -> Ys[k][j]
        for (int e = tid; e < TK * TJ; e += NT) {
            int kk, jj;
            if (p.ycs == 1) { jj = e % TJ; kk = e / TJ; } else { kk = e % TK; jj = e / TK; }
            T y = (T) 0;
            if (k0 + kk < p.K && j0 + jj < p.Q) y = p.Y[(k0 + kk) * p.yrs + (j0 + jj) * p.ycs];
            Ys[kk][jj] = y;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            T xa[4], yb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) xa[a] = Xs[kk][ti * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) yb[b] = Ys[kk][tj * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] += xa[a] * yb[b];
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t i = i0 + ti * 4 + a;
        if (i >= p.P) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t j = j0 + tj * 4 + b;
            if (j >= p.Q) continue;
            T* c = p.C + i * p.crs + j * p.ccs;
            T r = p.alpha * acc[a][b];
            if (p.beta != (T) 0) r += p.beta * (*c);
            *c = r;
        }
    }
}

template <typename T>
__global__ void scale_kernel(int64_t P, int64_t Q, T beta, T* __restrict__ C, int64_t crs, int64_t ccs,
                             int inner_is_j) {
    const int64_t total = P * Q;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t) gridDim.x * blockDim.x) {
        int64_t i, j;
        if (inner_is_j) { i = e / Q; j = e - i * Q; } else { j = e / P; i = e - j * P; }
        T* c = C + i * crs + j * ccs;
        *c = (beta == (T) 0) ? (T) 0 : beta * (*c);
    }
}

}  // namespace

template <typename T>
int launch_scale(int64_t P, int64_t Q, T beta, T* C, int64_t crs, int64_t ccs, cudaStream_t st) {
    if (P <= 0 || Q <= 0 || beta == (T) 1) return 0;
    int64_t total = P * Q;
    int64_t grid = (total + 255) / 256;
    int64_t cap = (int64_t) sm_count() * 16;
    if (grid > cap) grid = cap;
    scale_kernel<T><<<(unsigned) grid, 256, 0, st>>>(P, Q, beta, C, crs, ccs, ccs <= crs ? 1 : 0);
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}
// util::require_symmetric (RandBLAS/util.hh:128-148): the lexicographically first (i, j), i < j, with
// |A(i,j) - A(j,i)| > (|A(i,j)| + |A(j,i)| + 1) * tol, encoded as i * n + j (atomicMin); ~0 if the matrix is symmetric.
template <typename T>
__global__ void __launch_bounds__(256) symmetry_check_kernel(const T* __restrict__ A, int64_t n, int64_t rs, int64_t cs, T tol,
                                                             unsigned long long* __restrict__ first) {
    const int64_t total = n * n;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
        const int64_t i = e / n, j = e - i * n;
        if (j <= i) continue;
        const T aij = A[i * rs + j * cs], aji = A[j * rs + i * cs];
        const T viol = aij > aji ? aij - aji : aji - aij;
        const T rel = ((aij < 0 ? -aij : aij) + (aji < 0 ? -aji : aji) + (T) 1) * tol;
        if (viol > rel) atomicMin(first, (unsigned long long) e);
    }
}

template <typename T>
int launch_symmetry_check(const T* A, int64_t n, int64_t rs, int64_t cs, T tol, unsigned long long* first_dev, cudaStream_t st) {
    RB_CUDA(cudaMemsetAsync(first_dev, 0xff, 8, st));
    if (n > 1) {
        int64_t grid = (n * n + 255) / 256;
        const int64_t cap = (int64_t) sm_count() * 8;
        if (grid > cap) grid = cap;
        symmetry_check_kernel<T><<<(unsigned) grid, 256, 0, st>>>(A, n, rs, cs, tol, first_dev);
        count_launch();
        RB_CUDA(cudaGetLastError());
    }
    return 0;
}
template int launch_symmetry_check<float>(const float*, int64_t, int64_t, int64_t, float, unsigned long long*, cudaStream_t);
template int launch_symmetry_check<double>(const double*, int64_t, int64_t, int64_t, double, unsigned long long*, cudaStream_t);

template int launch_scale<float>(int64_t, int64_t, float, float*, int64_t, int64_t, cudaStream_t);
template int launch_scale<double>(int64_t, int64_t, double, double*, int64_t, int64_t, cudaStream_t);

template <typename T>
int launch_dense_generic(const DenseProblem<T>& p, cudaStream_t st) {
    if (p.P <= 0 || p.Q <= 0) return 0;
    if (p.K <= 0 || p.alpha == (T) 0) return launch_scale<T>(p.P, p.Q, p.beta, p.C, p.crs, p.ccs, st);
    dim3 grid((unsigned) ((p.Q + TJ - 1) / TJ), (unsigned) ((p.P + TI - 1) / TI));
    if (grid.y > 65535u) return fail("dense_generic: more than 65535 row tiles is not supported");
    const bool gauss = p.family == 'G';
    if (gauss && !p.S_buff && !p.gen.logtab) return fail_cuda(cudaErrorMemoryAllocation, "logf table of the Gaussian generator");
    if (p.S_buff) dense_generic_kernel<T, false, true><<<grid, NT, 0, st>>>(p);
    else if (gauss) dense_generic_kernel<T, true, false><<<grid, NT, 0, st>>>(p);
    else dense_generic_kernel<T, false, false><<<grid, NT, 0, st>>>(p);
    count_launch();
    RB_CUDA(cudaGetLastError());
    return 0;
}
template int launch_dense_generic<float>(const DenseProblem<float>&, cudaStream_t);
template int launch_dense_generic<double>(const DenseProblem<double>&, cudaStream_t);

}  // namespace rb
