// randblas_b200 -- header-only drop-in layer, part 1: enums, errors, RNGState.
//
// Same names, template parameters and semantics as the reference's RandBLAS/base.hh and exceptions.hh
// (reference file:line cited per item); every function that does work forwards to one extern "C" symbol of
// librandblas_b200.so (include/randblas_b200.h). Nothing here touches a GPU API directly.
#pragma once
#include <cstdint>
#include <cstring>
#include <exception>
#include <ostream>
#include <sstream>
#include <string>
#include <utility>
#include "../randblas_b200.h"

// BLAS++ is the reference's source of the Layout/Op enums (RandBLAS/base.hh:37). If the real <blas.hh> was
// included first its definitions are used; otherwise the two enums the hot path needs are provided with
// BLAS++'s underlying values.
#ifndef BLAS_HH
namespace blas {
enum class Layout : char { ColMajor = 'C', RowMajor = 'R' };
enum class Op : char { NoTrans = 'N', Trans = 'T', ConjTrans = 'C' };
}  // namespace blas
#endif

namespace RandBLAS {

// ---- RandBLAS/exceptions.hh:57-95,152-162 ----
class Error : public std::exception {
   public:
    Error() : std::exception() {}
    Error(std::string const& msg) : std::exception(), msg_(msg) {}
    Error(const char* msg, const char* func) : std::exception(), msg_(std::string(msg) + ", in function " + func) {}
    virtual const char* what() const noexcept override { return msg_.c_str(); }

   private:
    std::string msg_;
};

namespace internal {
inline void throw_if(bool cond, const char* condstr, const char* func) {
    if (cond) throw Error(condstr, func);
}
// a non-zero return of the C ABI becomes the exception the reference would have thrown
inline void check(int rc, const char* func) {
    if (rc != RB_OK) throw Error(rb_last_error(), func);
}
inline char to_char(blas::Layout l) { return l == blas::Layout::ColMajor ? 'C' : 'R'; }
inline char to_char(blas::Op o) { return o == blas::Op::NoTrans ? 'N' : 'T'; }
}  // namespace internal

#define randblas_error_if(cond) ::RandBLAS::internal::throw_if(cond, #cond, __func__)
#define randblas_require(cond) \
    ::RandBLAS::internal::throw_if(!(cond), "(" #cond ") was required, but did not hold", __func__)

}  // namespace RandBLAS

// ---- the slice of Random123 the API exposes through RNGState (array types with .v, operator[], incr) ----
namespace r123 {
template <int N>
struct Array32 {
    using value_type = uint32_t;
    static constexpr int static_size = N;
    uint32_t v[N];
    Array32() : v{} {}
    uint32_t& operator[](int i) { return v[i]; }
    const uint32_t& operator[](int i) const { return v[i]; }
    // little-endian multi-limb add (Random123 array.h `incr`; test_r123.cc:714-797)
    Array32& incr(unsigned long long n = 1) {
        unsigned long long carry = n;
        for (int i = 0; i < N && carry; ++i) {
            unsigned long long s = (unsigned long long) v[i] + (carry & 0xffffffffull);
            v[i] = (uint32_t) s;
            carry = (carry >> 32) + (s >> 32);
        }
        return *this;
    }
    bool operator==(const Array32& o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
    bool operator!=(const Array32& o) const { return !(*this == o); }
};
// Philox4x32-10: only the types are needed on the host; the block function runs on the device
// (rb_philox_words exposes it for bit-exact checks).
struct Philox4x32 {
    using ctr_type = Array32<4>;
    using key_type = Array32<2>;
};
}  // namespace r123

namespace RandBLAS {

using DefaultRNG = r123::Philox4x32;    // RandBLAS/base.hh:53

// ---- RandBLAS/base.hh:306-312 ----
enum class Axis : char { Short = 'S', Long = 'L' };

// ---- RandBLAS/base.hh:64-164 ----
template <typename RNG = DefaultRNG>
struct RNGState {
    using generator = RNG;
    using ctr_type = typename RNG::ctr_type;
    using key_type = typename RNG::key_type;
    using ctr_uint = typename RNG::ctr_type::value_type;
    using key_uint = typename RNG::key_type::value_type;
    static constexpr int len_c = RNG::ctr_type::static_size;
    static constexpr int len_k = RNG::key_type::static_size;
    static_assert(len_c == 4 && len_k == 2, "librandblas_b200 implements Philox4x32 states");

    ctr_type counter;
    key_type key;

    RNGState() : counter{}, key{} {}
    RNGState(uint64_t k) : counter{}, key{} { key.incr(k); }                 // base.hh:116-119
    RNGState(key_type const& k) : counter{}, key(k) {}
    RNGState(ctr_type const& c, key_type const& k) : counter(c), key(k) {}
    RNGState(const RNGState& s) = default;
    RNGState& operator=(const RNGState& s) = default;
    bool operator==(const RNGState& s) const { return counter == s.counter && key == s.key; }
    bool operator!=(const RNGState& s) const { return !(*this == s); }
};

// base.hh:172-189
template <typename RNG>
std::ostream& operator<<(std::ostream& out, const RNGState<RNG>& s) {
    out << "counter : {";
    for (int i = 0; i < s.len_c; ++i) out << s.counter[i] << (i + 1 < s.len_c ? ", " : "}\n");
    out << "key     : {";
    for (int i = 0; i < s.len_k; ++i) out << s.key[i] << (i + 1 < s.len_k ? ", " : "}");
    return out;
}

namespace internal {
template <typename RNG>
inline RNGState<RNG> with_counter(const RNGState<RNG>& s, const uint32_t ctr[4]) {
    RNGState<RNG> r(s);
    for (int i = 0; i < 4; ++i) r.counter.v[i] = ctr[i];
    return r;
}
}  // namespace internal

}  // namespace RandBLAS
