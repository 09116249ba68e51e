// Probe: tcgen05.mma kind::tf32 with the B operand MN-major (contiguous along N), 128B swizzle, written by hand in the
// canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (cute/atom/mma_traits_sm100.hpp, make_umma_desc<MN>).
// One CTA: D(128 x 256) = A(128 x 8, K-major SW128) * B(8 x 256). Variants (argv[1]):
//   0  K-major B (reference: the layout the library uses today)
//   1  MN-major B, 8-row atoms packed: MN block b at b * 1024 B (LBO = 1024), SBO = 8192 (next 8 k rows after all blocks)
//   2  MN-major B, TMA-box-like: MN block b at b * 4096 B (32 k rows of 128 B per box), LBO = 4096, SBO = 1024
//   3  as 2 with the LBO / SBO fields exchanged
//   4  MN-major B in the layout CUTLASS names as the only one for MN-major tf32 operands (sm100_common.inl:92):
//      SWIZZLE_128B_BASE32B (layout type 1), Swizzle<2,5,2>: atoms of 4 k rows x 128 B, the 32-byte chunk index XORed with
//      (k & 3); MN block b at b * 1024 B (LBO = 1024), the second 4-row atom of a K = 8 step 512 B further (SBO = 512)
//   5  as 4, TMA-box-like: MN block b at b * 4096 B (a box of 32 k rows), LBO = 4096, SBO = 512
// nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o mn_major_probe mn_major_probe.cu && ./mn_major_probe 1
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

constexpr int M = 128, N = 256, K = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int variant, uint32_t lbo, uint32_t sbo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* As = (float*) smem;                       // 128 rows x 128 B (K-major SW128), 16 KB
    float* Bs = (float*) (smem + 16384);             // up to 64 KB
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + 65536) / 4; i += 128) ((float*) smem)[i] = 0.f;
    __syncthreads();
    // A: row i, k: byte offset i * 128 + ((k / 4) ^ (i & 7)) * 16 + (k % 4) * 4
    for (int e = tid; e < M * K; e += 128) {
        const int i = e / K, k = e % K;
        As[(i * 128 + (((k >> 2) ^ (i & 7)) << 4) + (k & 3) * 4) / 4] = A[i * K + k];
    }
    for (int e = tid; e < K * N; e += 128) {
        const int k = e / N, j = e % N;
        size_t off;
        if (variant == 0) {                          // K-major: row j (128 B), k inside
            off = (size_t) j * 128 + (((k >> 2) ^ (j & 7)) << 4) + (k & 3) * 4;
        } else if (variant <= 3) {                   // MN-major: block b = j / 32, row k (128 B), chunk (j % 32) / 4
            const int b = j >> 5, c = (j & 31) >> 2;
            off = (size_t) b * lbo + (size_t) (k >> 3) * sbo + (size_t) (k & 7) * 128 + ((c ^ (k & 7)) << 4) + (j & 3) * 4;
        } else {                                     // MN-major, 128B swizzle with 32-byte atoms: 4-row groups
            const int b = j >> 5, c32 = (j & 31) >> 3;
            off = (size_t) b * lbo + (size_t) (k >> 2) * sbo + (size_t) (k & 3) * 128 + ((c32 ^ (k & 3)) << 5) + (j & 7) * 4;
        }
        Bs[off / 4] = B[k * N + j];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint64_t da = (uint64_t) ((smem_u32(As) >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t) (1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
        uint64_t db;
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
        if (variant == 0) {
            db = (uint64_t) ((smem_u32(Bs) >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t) (1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
        } else {
            const uint32_t f_lbo = (variant == 3) ? sbo : lbo, f_sbo = (variant == 3) ? lbo : sbo;
            db = (uint64_t) ((smem_u32(Bs) >> 4) & 0x3fffu) | ((uint64_t) ((f_lbo >> 4) & 0x3fffu) << 16) |
                 ((uint64_t) ((f_sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((variant >= 4 ? 1ull : 2ull) << 61);
            idesc |= (1u << 16);                     // B operand MN-major
        }
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the MMA
    {
        uint32_t ok = 0;
        for (int spin = 0; !ok && spin < (1 << 22); ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int cb = 0; cb < N / 32; ++cb) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t) (warp * 32) << 16) + (uint32_t) (cb * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 32; ++c) D[(size_t) tid * N + cb * 32 + c] = __uint_as_float(v[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    uint32_t lbo = 1024, sbo = 8192;
    if (variant >= 2) { lbo = 4096; sbo = 1024; }
    if (variant == 4) { lbo = 1024; sbo = 512; }
    if (variant == 5) { lbo = 4096; sbo = 512; }
    if (argc > 3) { lbo = (uint32_t) atoi(argv[2]); sbo = (uint32_t) atoi(argv[3]); }
    float *hA = new float[M * K], *hB = new float[K * N], *hD = new float[M * N];
    for (int i = 0; i < M * K; ++i) hA[i] = (float) ((i * 37) % 17 - 8);          // small integers: exact in tf32
    for (int i = 0; i < K * N; ++i) hB[i] = (float) ((i * 53) % 23 - 11);
    float *A, *B, *D;
    cudaMalloc(&A, M * K * 4); cudaMalloc(&B, K * N * 4); cudaMalloc(&D, M * N * 4);
    cudaMemcpy(A, hA, M * K * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(B, hB, K * N * 4, cudaMemcpyHostToDevice);
    cudaMemset(D, 0xff, M * N * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 65536);
    probe<<<1, 128, 16384 + 65536>>>(A, B, D, variant, lbo, sbo);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d (lbo %u sbo %u): CUDA error: %s\n", variant, lbo, sbo, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, D, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            double w = 0;
            for (int k = 0; k < K; ++k) w += (double) hA[i * K + k] * hB[k * N + j];
            const double err = fabs(w - hD[i * N + j]);
            if (!(err <= 1e-3)) ++bad;
            if (err > maxerr || err != err) maxerr = err;
        }
    printf("variant %d (lbo %u sbo %u): max |err| %.3g, %d of %d entries wrong\n", variant, lbo, sbo, maxerr, bad, M * N);
    return 0;
}
