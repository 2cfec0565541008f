// rt_fastdiv.hpp — division by an invariant 32-bit divisor; plain C++ so that the host (rt_api.cu fills the
// structure, tests/fastdiv_check.cpp checks it) and the device share one definition.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__
#else
#define RT_HD
#endif

namespace rtd {

// Division of a 32-bit unsigned number by an invariant divisor (Granlund & Montgomery; the branch-free form):
//   t = umulhi(m, x);  q = (t + ((x - t) >> s1)) >> s2        exact for every x < 2^32 and every d >= 1
// with l = ceil(log2 d), m = floor(2^32 (2^l - d) / d) + 1, s1 = min(l, 1), s2 = max(l - 1, 0).  The host fills it
// (make_fastdiv); the device spends 5 integer instructions instead of the ~20 (32-bit) / ~70 (64-bit) of a division.
struct FastDiv {
    uint32_t d, m, s1, s2;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f{d ? d : 1u, 0u, 0u, 0u};
    uint32_t l = 0;
    while (l < 32u && (1ull << l) < f.d) ++l;
    f.m = uint32_t(((1ull << 32) * ((1ull << l) - f.d)) / f.d + 1ull);
    f.s1 = l < 1u ? l : 1u;
    f.s2 = l > 1u ? l - 1u : 0u;
    return f;
}
RT_HD inline uint32_t fastdiv(uint32_t x, const FastDiv& f) {
#ifdef __CUDA_ARCH__
    const uint32_t t = __umulhi(f.m, x);
#else
    const uint32_t t = uint32_t((uint64_t(f.m) * x) >> 32);
#endif
    return (t + ((x - t) >> f.s1)) >> f.s2;
}

} // namespace rtd
