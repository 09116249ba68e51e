// Microbenchmark (not product code): throughput of reductions into an L2-resident slab of B as sketch_sparse issues
// them -- red.global.add.v4.f32 per lane vs cp.reduce.async.bulk (.add.f32) of 512 B / 2 KB from shared memory.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bulk red_bulk.cu ; run: ./red_bulk
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

// mode 0: red.v4 per lane, 128 floats per warp op. B column stride = ld floats, slab offset inside the column.
__global__ void k_red_v4(float* B, int ncols, int ld, int iters) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int it = 0; it < iters; ++it) {
        const uint32_t q = hash32(warp * 9781u + it) % (uint32_t) ncols;
        float* p = B + (size_t) q * ld + lane * 4;
        const float v = (float) it;
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
    }
}

// mode 1: FL floats per warp op through a ring of R shared-memory buffers and cp.reduce.async.bulk
template <int FL, int R>
__global__ void k_red_bulk(float* B, int ncols, int ld, int iters) {
    extern __shared__ __align__(128) float sm[];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float* ring = sm + (size_t) wl * R * FL;
    for (int it = 0; it < iters; ++it) {
        const uint32_t q = hash32(warp * 9781u + it) % (uint32_t) ncols;
        float* buf = ring + (it % R) * FL;
        if (it >= R) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(R - 1) : "memory");
            __syncwarp();
        }
        const float v = (float) it;
#pragma unroll
        for (int j = 0; j < FL / 128; ++j)
            *reinterpret_cast<float4*>(buf + j * 128 + lane * 4) = make_float4(v, v, v, v);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(B + (size_t) q * ld),
                         "r"(smem_u32(buf)), "n"(FL * 4)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
float timeit(F f, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const int ld = 512, iters = 2000;
    float* B;
    cudaMalloc(&B, (size_t) 125000 * ld * 4);
    cudaMemset(B, 0, (size_t) 125000 * ld * 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int ncols : {31250, 125000}) {
        for (int wpc : {8, 16}) {
            const int threads = wpc * 32, ctas = sms * (wpc == 8 ? 4 : 2);
            const double ops = (double) ctas * wpc * iters;
            float ms = timeit([&] { k_red_v4<<<ctas, threads>>>(B, ncols, ld, iters); }, 3);
            printf("ncols %6d warps/cta %2d  red.v4 x32 (512 B/warp op): %.3f ms  %.1f G red.v4/s  %.2f TB/s reduced\n", ncols, wpc, ms,
                   ops * 32 / ms / 1e6, ops * 512 / ms / 1e9);
            cudaFuncSetAttribute(k_red_bulk<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            cudaFuncSetAttribute(k_red_bulk<128, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            cudaFuncSetAttribute(k_red_bulk<512, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            ms = timeit([&] { k_red_bulk<128, 4><<<ctas, threads, wpc * 4 * 512>>>(B, ncols, ld, iters); }, 3);
            printf("ncols %6d warps/cta %2d  bulk 512 B ring 4: %.3f ms  %.2f TB/s reduced (= %.1f G red.v4/s equiv)\n", ncols, wpc, ms,
                   ops * 512 / ms / 1e9, ops * 32 / ms / 1e6);
            ms = timeit([&] { k_red_bulk<128, 8><<<ctas, threads, wpc * 8 * 512>>>(B, ncols, ld, iters); }, 3);
            printf("ncols %6d warps/cta %2d  bulk 512 B ring 8: %.3f ms  %.2f TB/s reduced (= %.1f G red.v4/s equiv)\n", ncols, wpc, ms,
                   ops * 512 / ms / 1e9, ops * 32 / ms / 1e6);
            if (wpc == 8) {
                ms = timeit([&] { k_red_bulk<512, 4><<<ctas, threads, wpc * 4 * 2048>>>(B, ncols, ld, iters / 4); }, 3);
                printf("ncols %6d warps/cta %2d  bulk 2 KB ring 4: %.3f ms  %.2f TB/s reduced (= %.1f G red.v4/s equiv)\n", ncols, wpc, ms,
                       ops / 4 * 2048 / ms / 1e9, ops / 4 * 128 / ms / 1e6);
            }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        }
    }
    return 0;
}
