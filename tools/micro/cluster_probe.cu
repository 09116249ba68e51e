// Probe: which cluster shapes does the driver accept for a 1-CTA-per-SM kernel (227 KB dynamic shared memory)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(576, 1) k(int* out) {
    extern __shared__ unsigned char sm[];
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.z == 0) out[blockIdx.y] = (int) r + (sm[0] & 0);
}
int main() {
    int* d; cudaMalloc(&d, 64 * 4); cudaMemset(d, 0xff, 64 * 4);
    const int smem = 220 * 1024;
    printf("setattr: %s\n", cudaGetErrorString(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)));
    for (int shape = 0; shape < 3; ++shape) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(4, 8, 83); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = shape == 0 ? 2 : 1; at[0].val.clusterDim.y = shape == 1 ? 2 : 1; at[0].val.clusterDim.z = 1;
        if (shape == 2) { at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 2; cfg.gridDim = dim3(8, 8, 2); }
        cfg.attrs = at; cfg.numAttrs = 1;
        int nclusters = -1;
        cudaError_t eo = cudaOccupancyMaxActiveClusters(&nclusters, k, &cfg);
        cudaError_t e = cudaLaunchKernelEx(&cfg, k, d);
        cudaError_t es = cudaDeviceSynchronize();
        int h[8]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("shape %d: occupancy %s clusters=%d launch %s sync %s ranks %d %d %d %d\n", shape, cudaGetErrorString(eo), nclusters,
               cudaGetErrorString(e), cudaGetErrorString(es), h[0], h[1], h[2], h[3]);
        cudaGetLastError();
    }
    return 0;
}
