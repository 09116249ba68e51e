// Micro-benchmark: what one DFMA of another warp costs a DMMA stream on the same scheduler (FP64 pipe sharing).
// 8 warps per CTA: warps 0-3 (one per scheduler) issue DMMAs back to back; warps 4-7 issue `nd` DFMAs per `gap` FFMAs,
// either as one dependent chain (ILP 1) or as `ilp` independent chains issued together (bursts).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_dfma dmma_dfma.cu && ./dmma_dfma
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, long long* cyc, int iters, int dfma_warps, int gap, volatile int* stop) {
    const int warp = threadIdx.x >> 5;
    __shared__ int done;
    if (threadIdx.x == 0) done = 0;
    __syncthreads();
    if (warp < 4) {
        double acc[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
        double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3 + 1.0;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
        }
        long long t1 = clock64();
        double s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
        if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
        __syncwarp();
        if ((threadIdx.x & 31) == 0) atomicAdd(&done, 1);
    } else if (warp < 4 + dfma_warps) {
        double x[ILP];
        float f = threadIdx.x;
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
        long long n = 0;
        while (*((volatile int*) &done) < 4) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(1.0000001), "d"(1e-9));
            for (int g = 0; g < gap; ++g) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(1e-3f));
            n += ILP;
        }
        double s = f;
#pragma unroll
        for (int i = 0; i < ILP; ++i) s += x[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
        if (threadIdx.x == 128 && blockIdx.x == 0) cyc[1] = n;
    }
}

template <int ILP>
void run(int dfma_warps, int gap) {
    double* out; long long* cyc; int* stop;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, 16);
    cudaMalloc(&stop, 4);
    cudaMemset(cyc, 0, 16);
    const int iters = 4000;
    k<ILP><<<148, 256>>>(out, cyc, iters, dfma_warps, gap, stop);
    cudaDeviceSynchronize();
    long long c[2];
    cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost);
    const double ndmma = iters * 8.0;
    printf("DFMA warps/scheduler %d, ilp %d, %3d FFMA between groups: %.2f cycles per DMMA; %lld DFMA issued by one warp = %.3f per DMMA; "
           "extra cycles per DFMA %.2f\n", dfma_warps / 4, ILP, gap, c[0] / ndmma, c[1], c[1] / ndmma,
           c[1] ? (c[0] - 16.0 * ndmma) / ((double) c[1] * (dfma_warps / 4)) : 0.0);
    cudaFree(out); cudaFree(cyc); cudaFree(stop);
}

int main() {
    run<1>(0, 0);
    for (int gap : {0, 8, 32, 128}) { run<1>(4, gap); run<4>(4, gap * 4); run<8>(4, gap * 8); }
    return 0;
}
