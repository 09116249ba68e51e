// TEST INFRASTRUCTURE. Exhaustive host check (all 2^32 Philox words, ~15 s on 8 cores) that the evaluation scheme the
// device uses for the Box-Muller angle (randblas_b200/csrc/philox.cuh: quadrant from two comparisons on uneg11(w),
// float->double by bit manipulation, glibc 2.39 __sincosf_fma polynomial in FMA arithmetic) reproduces this host's
// sincosf bit for bit. Build: gcc -O2 -march=x86-64-v3 -fopenmp -ffp-contract=off validate_libm_model.c -lm
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <omp.h>
static inline float u2f(uint32_t u){float f; memcpy(&f,&u,4); return f;}
static inline uint32_t f2u(float f){uint32_t u; memcpy(&u,&f,4); return u;}
static inline double u2d(uint64_t u){double f; memcpy(&f,&u,8); return f;}
static inline uint64_t d2u(double f){uint64_t u; memcpy(&u,&f,8); return u;}
static const double hpi_inv = 0x1.45f306dc9c883p+23, nhpi = -0x1.921fb54442d18p+0;
static const double C0=1.0,C1=-0x1.ffffffd0c621cp-2,C2=0x1.55553e1068f19p-5,C3=-0x1.6c087e89a359dp-10,C4=0x1.99343027bf8c3p-16;
static const double S1=-0x1.555545995a603p-3,S2=0x1.1107605230bc4p-7,S3=-0x1.994eb3774cf24p-13;
static inline float uneg11(uint32_t w){ return fmaf((float)(int32_t)w, 0x1p-31f, 0x1p-32f); }
// proposed device model: quadrant from comparisons on x, theta->double by bit manipulation
static inline void model_sincos(uint32_t w, float* sn, float* cs){
    float x = uneg11(w);
    float th = 3.1415926535897932f * x;
    float ax = fabsf(x);
    int na = (ax >= 0.25f) + (ax >= 0.75f);
    int n = (x < 0) ? -na : na;
    // float -> double by bits (normal numbers only)
    uint32_t u = f2u(th), ua = u & 0x7fffffffu;
    uint64_t wide = (uint64_t)ua * 0x20000000ull + 0x3800000000000000ull;
    wide |= (uint64_t)(u & 0x80000000u) << 32;
    double xd = u2d(wide);
    if (xd != (double)th) { *sn = NAN; *cs = NAN; return; }
    double xr = fma((double)n, nhpi, xd);
    double xs = ((n + 1) & 2) ? -xr : xr;
    double x2 = xr*xr, x3 = x2*xs, x4 = x2*x2;
    double s1 = fma(x2,S3,S2), c2 = fma(x2,C4,C3), c1 = fma(x2,C1,C0);
    double x5 = x2*x3, x6 = x2*x4;
    double s = fma(x3,S1,xs), c = fma(x4,C2,c1);
    s = fma(s1,x5,s); c = fma(c2,x6,c);
    float fs = (float)s, fc = (float)c;
    if (n & 2) fc = -fc;
    if (n & 1) { *sn = fc; *cs = fs; } else { *sn = fs; *cs = fc; }
}
int main(){
    long bad = 0, badn = 0;
    #pragma omp parallel for reduction(+:bad,badn) schedule(static)
    for (long long wi = 0; wi < (1LL<<32); ++wi) {
        uint32_t w = (uint32_t)wi;
        float x = uneg11(w); float th = 3.1415926535897932f * x;
        float s0, c0; sincosf(th, &s0, &c0);
        float s1, c1; model_sincos(w, &s1, &c1);
        if (f2u(s0)!=f2u(s1) || f2u(c0)!=f2u(c1)) bad++;
        // quadrant check vs glibc's reduce_fast
        int ng = ((int32_t)((double)th * hpi_inv) + 0x800000) >> 24;
        float ax = fabsf(x); int na = (ax >= 0.25f) + (ax >= 0.75f); int n = (x<0)?-na:na;
        if (n != ng) badn++;
    }
    printf("sincos mismatches: %ld, quadrant mismatches: %ld\n", bad, badn);
    return 0;
}
