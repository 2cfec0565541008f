// Host check of rtd::FastDiv (csrc/rt_fastdiv.hpp): exact quotients for every divisor class the renderer uses (frame
// widths, pixel counts < 2^31) over edge values and a random sweep.  Built and run by tests/test_host_api.py.
#include <cstdint>
#include <cstdio>
#include <random>

#include "../raytracing_renderer_cuda_b200/csrc/rt_fastdiv.hpp"

int main() {
    std::mt19937_64 g(7);
    const uint32_t ds[] = {1u, 2u, 3u, 5u, 7u, 37u, 600u, 1200u, 1920u, 3840u, 7680u, 720000u, 2073600u, 8294400u, 33177600u,
                           (1u << 16), (1u << 16) + 1u, (1u << 24) - 1u, (1u << 30), (1u << 31) - 1u, (1u << 31), 0xfffffffeu, 0xffffffffu};
    unsigned long long checked = 0;
    auto check = [&](uint32_t d, uint32_t x) {
        const rtd::FastDiv f = rtd::make_fastdiv(d);
        if (rtd::fastdiv(x, f) != x / d) {
            printf("FAIL d=%u x=%u got=%u want=%u\n", d, x, rtd::fastdiv(x, f), x / d);
            return false;
        }
        ++checked;
        return true;
    };
    for (uint32_t d : ds) {
        const uint32_t xs[] = {0u, 1u, d - 1u, d, d + 1u, 2u * d - 1u, 2u * d, 0x7fffffffu, 0x80000000u, 0xfffffffeu, 0xffffffffu};
        for (uint32_t x : xs)
            if (!check(d, x)) return 1;
        for (int k = 0; k < 200000; ++k)
            if (!check(d, uint32_t(g()))) return 1;
    }
    for (int k = 0; k < 2000000; ++k) {
        uint32_t d = uint32_t(g() >> (g() & 31));
        if (d == 0) d = 1;
        if (!check(d, uint32_t(g()))) return 1;
        if (!check(d, d * uint32_t(g() & 15) + uint32_t(g() % d))) return 1;
    }
    printf("ok %llu\n", checked);
    return 0;
}
