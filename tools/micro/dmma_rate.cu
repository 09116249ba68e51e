// Micro-benchmark: issue rate of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) per scheduler as a function of the number of
// warps per scheduler and of the number of independent accumulators per warp. One CTA per SM, W warps per scheduler.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_rate dmma_rate.cu && ./dmma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k(double* out, long long* cyc, int iters) {
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3 + 1.0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NACC>
void run(int warps_per_sched) {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<NACC><<<148, 128 * warps_per_sched>>>(out, cyc, iters);
    k<NACC><<<148, 128 * warps_per_sched>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_sched = (double) c / ((double) iters * NACC * warps_per_sched);
    printf("warps/scheduler %d, independent accumulators %2d: %.2f cycles per DMMA per scheduler (%.2f per warp)\n",
           warps_per_sched, NACC, per_sched, (double) c / ((double) iters * NACC));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w = 1; w <= 4; ++w) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); run<32>(w); }
    return 0;
}
