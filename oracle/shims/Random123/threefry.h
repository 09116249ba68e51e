// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Threefry4x32-R restated from the published algorithm; the reference library never
// calls it on the hot path -- it is here only so the uneg11/u01 histograms of
// test/test_basic_rng/test_r123.cc:568-671 can be reproduced as a pin for uniform.hpp.
// Pinned by r123_kat_vectors.txt:47-55.
#pragma once
#include "Random123/array.h"

namespace r123 {
template <unsigned R>
struct Threefry4x32_R {
    typedef r123array4x32 ctr_type;
    typedef r123array4x32 key_type;
    typedef r123array4x32 ukey_type;
    static inline uint32_t rotl(uint32_t x, unsigned n) { return (x << (n & 31)) | (x >> ((32 - n) & 31)); }
    ctr_type operator()(ctr_type in, key_type k) const {
        static const unsigned R0[8] = {10, 11, 13, 23, 6, 17, 25, 18};
        static const unsigned R1[8] = {26, 21, 27, 5, 20, 11, 10, 20};
        uint32_t ks[5];
        ks[4] = 0x1BD11BDAu;
        for (int i = 0; i < 4; ++i) { ks[i] = k.v[i]; ks[4] ^= k.v[i]; }
        uint32_t X0 = in.v[0] + ks[0], X1 = in.v[1] + ks[1], X2 = in.v[2] + ks[2], X3 = in.v[3] + ks[3];
        for (unsigned r = 0; r < R; ++r) {
            if ((r & 1) == 0) {
                X0 += X1; X1 = rotl(X1, R0[r % 8]); X1 ^= X0;
                X2 += X3; X3 = rotl(X3, R1[r % 8]); X3 ^= X2;
            } else {
                X0 += X3; X3 = rotl(X3, R0[r % 8]); X3 ^= X0;
                X2 += X1; X1 = rotl(X1, R1[r % 8]); X1 ^= X2;
            }
            if ((r % 4) == 3) {
                unsigned s = r / 4 + 1;
                X0 += ks[s % 5]; X1 += ks[(s + 1) % 5]; X2 += ks[(s + 2) % 5]; X3 += ks[(s + 3) % 5];
                X3 += s;
            }
        }
        ctr_type o; o.v[0] = X0; o.v[1] = X1; o.v[2] = X2; o.v[3] = X3;
        return o;
    }
};
typedef Threefry4x32_R<20> Threefry4x32;
} // namespace r123
