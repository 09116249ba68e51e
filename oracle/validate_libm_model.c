// TEST INFRASTRUCTURE. Exhaustive host check (all 2^32 Philox words, ~1 min on 8 cores) that the evaluation schemes
// the device uses for Box-Muller (randblas_b200/csrc/philox.cuh) reproduce THIS host's libm bit for bit:
//   angle : uneg11(w) -> theta = fl32(pi_f * x) -> quadrant n = rint(2x) by the magic-number add -> reduced argument by
//           one exact-table add -> glibc 2.39 __sincosf_fma's polynomials in Horner form (FMA arithmetic) -> float
//   radius: u01(w) -> glibc __logf_fma with the (k, i) -> {invc, logc + k*ln2} table folded into one lookup and the
//           float->double conversion done by bit manipulation -> float
// Build: gcc -O2 -march=x86-64-v3 -fopenmp -ffp-contract=off validate_libm_model.c -lm && ./a.out
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>
static inline uint32_t f2u(float f){uint32_t u; memcpy(&u,&f,4); return u;}
static inline double u2d(uint64_t u){double f; memcpy(&f,&u,8); return f;}
static const double C0=1.0,C1=-0x1.ffffffd0c621cp-2,C2=0x1.55553e1068f19p-5,C3=-0x1.6c087e89a359dp-10,C4=0x1.99343027bf8c3p-16;
static const double S1=-0x1.555545995a603p-3,S2=0x1.1107605230bc4p-7,S3=-0x1.994eb3774cf24p-13;
static const double LN2 = 0x1.62e42fefa39efp-1, A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
static const double LOGTAB[32] = {
    0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2,
    0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3,
    0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4,
    0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1.0000000000000p+0, 0x0.0p+0,
    0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5,  0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4,
    0x1.b2036576afce6p-1, 0x1.526e57720db08p-3,  0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3,
    0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,  0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2,
};
static double BIG[544][2];
static inline float uneg11(uint32_t w){ return fmaf((float)(int32_t)w, 0x1p-31f, 0x1p-32f); }
static inline float u01(uint32_t w){ return fmaf((float)w, 0x1p-32f, 0x1p-33f); }
// the device shifts the float's bits into double position WITHOUT re-biasing the exponent (value * 2^-896) and
// compensates in the operand it multiplies with: invc * 2^896 in the table, fma(xd', 2^896, t) for the angle
static inline double f2d_pos_unbiased(uint32_t u){ return u2d((uint64_t)u * 0x20000000ull); }

static const double QT[5] = {0x1.921fb54442d18p+1, 0x1.921fb54442d18p+0, 0.0, -0x1.921fb54442d18p+0, -0x1.921fb54442d18p+1};  // -n pi/2, n = -2..2
static inline void model_sincos(uint32_t w, float* sn, float* cs){
    const float x = uneg11(w);
    const float th = 3.1415926535897932f * x;
    // quadrant n = round-to-nearest-even(2x) read from the low mantissa bits of 2x + 1.5*2^23
    const uint32_t tb = f2u(fmaf(x, 2.0f, 12582912.0f));
    const int n = (int)(tb - 0x4B400000u);
    const uint32_t u = f2u(th);
    uint64_t xb = (uint64_t)(u & 0x7fffffffu) * 0x20000000ull;
    xb |= (uint64_t)(u & 0x80000000u) << 32;
    const double xd = u2d(xb);                      // theta * 2^-896
    const double xr = fma(xd, 0x1p896, QT[n + 2]);
    const uint32_t sinflip = (tb * 0x40000000u + 0x40000000u) & 0x80000000u;   // (n + 1) & 2
    const uint32_t cosflip = (tb * 0x40000000u) & 0x80000000u;                 // n & 2
    const int swap = tb & 1;                                                   // n & 1
    const double xs = sinflip ? -xr : xr;
    const double x2 = xr * xr, x3 = x2 * xs;
    double p = fma(S3, x2, S2); p = fma(p, x2, S1);
    const double s = fma(x3, p, xs);
    double c = fma(C4, x2, C3); c = fma(c, x2, C2); c = fma(c, x2, C1); c = fma(c, x2, C0);
    const float fs = (float)s;
    float fc = (float)c;
    if (cosflip) fc = -fc;
    *sn = swap ? fc : fs;
    *cs = swap ? fs : fc;
}
static inline float model_log(uint32_t w){
    const float x = u01(w);
    const uint32_t ix = f2u(x);
    const int32_t tmp = (int32_t)(ix - 0x3f330000u);
    const int idx = (tmp >> 19) + 528;
    const uint32_t iz = ix - ((uint32_t)tmp & 0xff800000u);
    const double z = f2d_pos_unbiased(iz);          // z * 2^-896
    const double r = fma(z, BIG[idx][0], -1.0);
    const double r2 = r * r;
    double y = fma(A1, r, A2);
    y = fma(A0, r2, y);
    y = fma(y, r2, BIG[idx][1] + r);
    return (float)y;
}
int main(void){
    for (int k = -33; k <= 0; ++k)
        for (int i = 0; i < 16; ++i) {
            BIG[(k + 33) * 16 + i][0] = ldexp(LOGTAB[2 * i], 896);
            BIG[(k + 33) * 16 + i][1] = fma((double)k, LN2, LOGTAB[2 * i + 1]);
        }
    long bad_sc = 0, bad_log = 0, bad_idx = 0;
    #pragma omp parallel for reduction(+:bad_sc,bad_log,bad_idx) schedule(static)
    for (long long wi = 0; wi < (1LL<<32); ++wi) {
        const uint32_t w = (uint32_t)wi;
        float s0, c0, s1, c1;
        sincosf(3.1415926535897932f * uneg11(w), &s0, &c0);
        model_sincos(w, &s1, &c1);
        if (f2u(s0) != f2u(s1) || f2u(c0) != f2u(c1)) bad_sc++;
        const float x = u01(w);
        const int32_t tmp = (int32_t)(f2u(x) - 0x3f330000u);
        const int idx = (tmp >> 19) + 528;
        if (idx < 0 || idx >= 544) { bad_idx++; continue; }
        if (f2u(logf(x)) != f2u(model_log(w))) bad_log++;
    }
    printf("sincos mismatches: %ld, logf mismatches: %ld, table index out of range: %ld (of 2^32 words each)\n", bad_sc,
           bad_log, bad_idx);
    return (bad_sc || bad_log || bad_idx) ? 1 : 0;
}
