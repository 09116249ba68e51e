// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Fixed-width counter/key arrays with the extended-precision `incr` semantics the
// reference relies on (RandBLAS/dense_skops.hh:129,141,154,168; base.hh:119) and that
// test/test_basic_rng/test_r123.cc:710-797 pins (little-endian limbs, carry across limbs).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include "Random123/features/compilerfeatures.h"

namespace r123 {
template <typename W, int N>
struct Array {
    typedef W value_type;
    enum { static_size = N };
    W v[N];
    W& operator[](size_t i) { return v[i]; }
    const W& operator[](size_t i) const { return v[i]; }
    W* data() { return v; }
    const W* data() const { return v; }
    W* begin() { return v; }
    W* end() { return v + N; }
    const W* begin() const { return v; }
    const W* end() const { return v + N; }
    static size_t size() { return N; }
    bool operator==(const Array& o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
    bool operator!=(const Array& o) const { return !(*this == o); }
    // add 1 with carry propagating up the limbs
    Array& incr() {
        for (int i = 0; i < N; ++i) { if (++v[i] != 0) break; }
        return *this;
    }
    // add a 64-bit integer; limbs are little-endian; carry propagates across limbs
    Array& incr(unsigned long long n) {
        if constexpr (sizeof(W) == 8) {
            unsigned long long carry = n;
            for (int i = 0; i < N && carry != 0; ++i) {
                unsigned long long old = v[i];
                v[i] = (W)(old + carry);
                carry = (v[i] < old) ? 1ull : 0ull;
            }
        } else {
            unsigned long long carry = 0;
            for (int i = 0; i < N; ++i) {
                unsigned long long s = (unsigned long long) v[i] + (n & 0xffffffffull) + carry;
                v[i] = (W) s;
                carry = s >> 32;
                n >>= 32;
                if (n == 0 && carry == 0) break;
            }
        }
        return *this;
    }
};
} // namespace r123

typedef r123::Array<uint32_t, 2> r123array2x32;
typedef r123::Array<uint32_t, 4> r123array4x32;
typedef r123::Array<uint64_t, 2> r123array2x64;
typedef r123::Array<uint64_t, 4> r123array4x64;
