// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Minimal stand-in for Random123's compiler-feature header; Random123 itself is an
// un-vendored third-party dependency of the reference.
#pragma once
#include <cstdint>
#include <cassert>
#ifndef R123_ASSERT
#define R123_ASSERT(x) assert(x)
#endif
#define R123_USE_64BIT 1
#define R123_STATIC_INLINE static inline
#define R123_CUDA_DEVICE
#define R123_CONSTEXPR constexpr
