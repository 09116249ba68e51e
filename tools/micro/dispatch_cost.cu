// Micro-benchmark: how many dispatch cycles does one warp instruction of a given kind take from its scheduler, and does it
// hold the dispatch port against other warps' FFMAs? (fill_dense Gaussian: issue slots 66% busy with "dispatch_stall" as
// the top reason and no pipe above 50%.) One CTA of 1024 threads per SM = 8 warps per scheduler; every warp runs
// iterations of NX independent instructions of kind X followed by NF independent FFMAs. Reported: cycles per iteration and
// scheduler divided by the 8 warps, i.e. the dispatch cycles one warp's iteration costs. NX = 0 calibrates FFMA at 1.0.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dispatch_cost dispatch_cost.cu && ./dispatch_cost
#include <cstdio>
#include <cuda_runtime.h>

enum { K_NONE, K_DFMA, K_IMADW, K_F2D, K_D2F, K_RSQ, K_LOP3, K_I2F, K_IMAD, K_F2I_D, K_I2D, K_CVTPAIR, K_I2F_F2I, K_I2D_D2I };
static const char* names[] = {"none", "DFMA", "IMAD.WIDE.U32", "F2F.F64.F32", "F2F.F32.F64", "MUFU.RSQ", "LOP3", "I2FP.F32.S32",
                              "IMAD (32-bit)", "F2I.S32.F64", "I2F.F64.S32", "F2F.F64.F32+F2F.F32.F64", "I2FP.F32.S32+F2I.S32.F32", "I2F.F64.S32+F2I.S32.F64"};

template <int KIND, int NX, int NF, int PART = 0>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, int iters) {
    float f[8];
    double d[8];
    unsigned u[8];
    unsigned long long w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        f[i] = threadIdx.x * 1e-3f + i;
        d[i] = threadIdx.x * 1e-3 + i;
        u[i] = threadIdx.x * 2654435761u + i;
        w[i] = u[i];
    }
    float g[8];
    unsigned v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = threadIdx.x + i; v[i] = threadIdx.x * 40503u + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            const int i = j & 7;
            if (KIND == K_DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(1.0000001), "d"(1e-9));
            if (KIND == K_IMADW) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(u[i]), "r"(0xD2511F53u));
            if (KIND == K_F2D) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i]));
            if (KIND == K_D2F) asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(d[i]));
            if (KIND == K_RSQ) asm volatile("rsqrt.approx.f32 %0, %0;" : "+f"(f[i]));
            if (KIND == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(0x9E3779B9u));
            if (KIND == K_I2F) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[i]) : "r"(u[i]));
            if (KIND == K_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(0xD2511F53u), "r"(12345u));
            if (KIND == K_F2I_D) asm volatile("cvt.rni.s32.f64 %0, %1;" : "=r"(u[i]) : "d"(d[i]));
            if (KIND == K_I2D) asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d[i]) : "r"(u[i]));
            // dependent chains (a conversion whose result is dead is removed by ptxas): X = the PAIR of conversions
            if (KIND == K_CVTPAIR) asm volatile("{\n\t.reg .f64 t;\n\tcvt.f64.f32 t, %0;\n\tcvt.rn.f32.f64 %0, t;\n\t}" : "+f"(f[i]));
            if (KIND == K_I2F_F2I) asm volatile("{\n\t.reg .f32 t;\n\tcvt.rn.f32.s32 t, %0;\n\tcvt.rzi.s32.f32 %0, t;\n\t}" : "+r"(u[i]));
            if (KIND == K_I2D_D2I) asm volatile("{\n\t.reg .f64 t;\n\tcvt.rn.f64.s32 t, %0;\n\tcvt.rzi.s32.f64 %0, t;\n\t}" : "+r"(u[i]));
        }
#pragma unroll
        for (int j = 0; j < NF; ++j) {
            if (PART == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(g[j & 7]) : "f"(1.0001f), "f"(1e-3f));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[j & 7]) : "r"(v[(j + 1) & 7]), "r"(0x9E3779B9u));
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i] + (float) d[i] + (float) u[i] + (float) w[i] + g[i] + (float) v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND, int NX, int NF, int PART = 0>
void run(float* out, long long* cyc) {
    const int iters = 2000;
    k<KIND, NX, NF, PART><<<148, 1024>>>(out, cyc, iters);
    k<KIND, NX, NF, PART><<<148, 1024>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0;
    for (int i = 0; i < 148; ++i) m += (double) h[i];
    m /= 148;
    const double per = m / iters / 8.0;      // cycles of the scheduler per warp-iteration
    const double x = NX ? (per - NF) / NX : 0.0;
    printf("%-26s partner %s NX=%2d NF=%2d : %7.2f cycles per warp-iteration  -> %5.2f per X if the partner costs 1.0 (sum of issue slots %d)\n",
           names[KIND], PART ? "LOP3" : "FFMA", NX, NF, per, x, NX + NF);
}

template <int KIND>
void sweep(float* out, long long* cyc) {
    run<KIND, 8, 0>(out, cyc);
    run<KIND, 8, 8>(out, cyc);
    run<KIND, 8, 24>(out, cyc);
    run<KIND, 4, 28>(out, cyc);
    run<KIND, 2, 30>(out, cyc);
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    run<K_NONE, 0, 32>(out, cyc);
    sweep<K_DFMA>(out, cyc);
    sweep<K_IMADW>(out, cyc);
    sweep<K_IMAD>(out, cyc);
    sweep<K_RSQ>(out, cyc);
    sweep<K_LOP3>(out, cyc);
    sweep<K_CVTPAIR>(out, cyc);
    sweep<K_I2F_F2I>(out, cyc);
    sweep<K_I2D_D2I>(out, cyc);
    // the same with LOP3 (16-lane integer ALU pipe) as the partner instruction
    run<K_NONE, 0, 32, 1>(out, cyc);
    run<K_IMADW, 8, 24, 1>(out, cyc);
    run<K_IMADW, 4, 28, 1>(out, cyc);
    run<K_DFMA, 8, 24, 1>(out, cyc);
    run<K_RSQ, 2, 30, 1>(out, cyc);
    run<K_CVTPAIR, 2, 30, 1>(out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
