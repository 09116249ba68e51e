// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Box-Muller restated from Random123's published boxmuller.hpp (host, non-CUDA branch):
//   sincosf(PIf * uneg11<float>(u0)), r = sqrtf(-2.f * logf(u01<float>(u1))).
// Called by the reference at RandBLAS/random_gen.hh:69. No golden Gaussian values exist
// in the reference tree (only KS tests, test_continuous.cc:146-167): at bit level the
// Gaussian path is "parity unpinned" with respect to upstream Random123; it is pinned
// against this host libm (glibc) only.
#pragma once
#include <cmath>
#include <cstdint>
#include "Random123/uniform.hpp"

namespace r123 {
struct float2  { float x, y; };
struct double2 { double x, y; };

static inline float2 boxmuller(uint32_t u0, uint32_t u1) {
    const float PIf = 3.1415926535897932f;
    float2 f;
    ::sincosf(PIf * uneg11<float>(u0), &f.x, &f.y);
    float r = ::sqrtf(-2.f * ::logf(u01<float>(u1)));  // u01 never returns 0
    f.x *= r;
    f.y *= r;
    return f;
}

static inline double2 boxmuller(uint64_t u0, uint64_t u1) {
    const double PI = 3.1415926535897932;
    double2 f;
    ::sincos(PI * uneg11<double>(u0), &f.x, &f.y);
    double r = ::sqrt(-2. * ::log(u01<double>(u1)));
    f.x *= r;
    f.y *= r;
    return f;
}
} // namespace r123
