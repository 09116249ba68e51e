// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// u01 / uneg11 restated from Random123's published uniform.hpp:
//   u01<F>(w)    = F(unsigned(w)) * 2^-W + 2^-(W+1)
//   uneg11<F>(w) = F(signed(w))   * 2^-(W-1) + 2^-W
// where the constants are computed in F exactly as Random123 does
// (factor = 1/(F(max)+1), halffactor = 0.5*factor). Called by the reference at
// RandBLAS/random_gen.hh:135 and RandBLAS/sparse_data/random_matrix.hh:89.
// Pin: the 26-bin histograms of test/test_basic_rng/test_r123.cc:612,618.
#pragma once
#include <array>
#include <limits>
#include <type_traits>
#include "Random123/features/compilerfeatures.h"

namespace r123 {
template <typename T> struct make_signed   { typedef typename std::make_signed<T>::type type; };
template <typename T> struct make_unsigned { typedef typename std::make_unsigned<T>::type type; };
template <typename T> constexpr T maxTvalue() { return std::numeric_limits<T>::max(); }

template <typename Ftype, typename Itype>
static inline Ftype u01(Itype in) {
    typedef typename make_unsigned<Itype>::type Utype;
    constexpr Ftype factor = Ftype(1.) / (Ftype(maxTvalue<Utype>()) + Ftype(1.));
    constexpr Ftype halffactor = Ftype(0.5) * factor;
    return Utype(in) * factor + halffactor;
}

template <typename Ftype, typename Itype>
static inline Ftype uneg11(Itype in) {
    typedef typename make_signed<Itype>::type Stype;
    constexpr Ftype factor = Ftype(1.) / (Ftype(maxTvalue<Stype>()) + Ftype(1.));
    constexpr Ftype halffactor = Ftype(0.5) * factor;
    return Stype(in) * factor + halffactor;
}

template <typename Ftype, typename Itype>
static inline Ftype u01fixedpt(Itype in) {
    typedef typename make_unsigned<Itype>::type Utype;
    constexpr int excess = std::numeric_limits<Utype>::digits - std::numeric_limits<Ftype>::digits;
    if (excess >= 0) {
        constexpr int ex_nowarn = (excess >= 0) ? excess : 0;
        constexpr Ftype factor = Ftype(1.) / (Ftype(1.) + Ftype((maxTvalue<Utype>() >> ex_nowarn)));
        return (1 | (Utype(in) >> ex_nowarn)) * factor;
    } else {
        return u01<Ftype>(in);
    }
}

template <typename Ftype, typename CollType>
static inline std::array<Ftype, CollType::static_size> u01all(CollType in) {
    std::array<Ftype, CollType::static_size> ret;
    for (size_t i = 0; i < (size_t) CollType::static_size; ++i) ret[i] = u01<Ftype>(in[i]);
    return ret;
}
template <typename Ftype, typename CollType>
static inline std::array<Ftype, CollType::static_size> uneg11all(CollType in) {
    std::array<Ftype, CollType::static_size> ret;
    for (size_t i = 0; i < (size_t) CollType::static_size; ++i) ret[i] = uneg11<Ftype>(in[i]);
    return ret;
}
} // namespace r123
