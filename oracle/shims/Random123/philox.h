// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Philox4x32-R restated from the published algorithm (Salmon et al., SC'11; Random123
// philox.h). Pinned by the reference's known-answer vectors
// test/test_basic_rng/r123_kat_vectors.txt:16-21 (see tests/test_oracle_kat.py).
#pragma once
#include "Random123/array.h"

namespace r123 {
template <unsigned R>
struct Philox4x32_R {
    typedef r123array4x32 ctr_type;
    typedef r123array2x32 key_type;
    typedef r123array2x32 ukey_type;
    static const unsigned rounds = R;
    ctr_type operator()(ctr_type c, key_type k) const {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
        for (unsigned r = 0; r < R; ++r) {
            if (r > 0) { k.v[0] += W0; k.v[1] += W1; }
            uint64_t p0 = (uint64_t) M0 * c.v[0];
            uint64_t p1 = (uint64_t) M1 * c.v[2];
            ctr_type o;
            o.v[0] = (uint32_t)(p1 >> 32) ^ c.v[1] ^ k.v[0];
            o.v[1] = (uint32_t) p1;
            o.v[2] = (uint32_t)(p0 >> 32) ^ c.v[3] ^ k.v[1];
            o.v[3] = (uint32_t) p0;
            c = o;
        }
        return c;
    }
};
typedef Philox4x32_R<10> Philox4x32;
} // namespace r123
