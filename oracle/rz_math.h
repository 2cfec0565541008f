// oracle/rz_math.h — TEST INFRASTRUCTURE.  Exact host emulation of the CUDA round-toward-zero
// intrinsics the reference's vec3 is written in (vec3.h:5-7,73-151,258-347): __fadd_rz,
// __fsub_rz, __fmul_rz, __fdiv_rz, __fsqrt_rz.  x86-64 only: one asm statement switches MXCSR
// to round-toward-zero, issues the scalar SSE instruction and restores MXCSR, so the compiler
// cannot move the operation out of the rounding-mode window and other threads are unaffected.
#pragma once

#if !defined(__x86_64__)
#error "oracle/rz_math.h emulates CUDA's _rz intrinsics with SSE rounding control; x86-64 only"
#endif

#include <cmath>
#include <cstdint>

namespace rz {

#ifdef ORACLE_RZ_AS_RN
// Timing build only (oracle/_ref/libref_cpu_rn.so, bench.py's host-core baseline): the rounding-mode
// switches of the exact emulation cost more than the arithmetic, which would flatter the GPU.  Here the
// _rz intrinsics round to nearest — SURVEY.md Appendix B's shim — so the CPU runs at its natural speed.
static inline float add(float a, float b) { return a + b; }
static inline float sub(float a, float b) { return a - b; }
static inline float mul(float a, float b) { return a * b; }
static inline float div(float a, float b) { return a / b; }
static inline float sqrt(float a) { return ::sqrtf(a); }
static inline float fma(float a, float b, float c) { return float(double(a) * double(b) + double(c)); }
#else
#define RZ_BINOP(name, insn)                                                                  \
    static inline float name(float a, float b) {                                               \
        uint32_t saved, mode;                                                                  \
        __asm__ volatile("stmxcsr %0" : "=m"(saved));                                          \
        mode = saved | 0x6000u; /* RC = 11b: truncate */                                       \
        __asm__ volatile("ldmxcsr %2\n\t" insn " %1, %0\n\tldmxcsr %3"                         \
                         : "+x"(a)                                                             \
                         : "x"(b), "m"(mode), "m"(saved));                                     \
        return a;                                                                              \
    }
RZ_BINOP(add, "addss")
RZ_BINOP(sub, "subss")
RZ_BINOP(mul, "mulss")
RZ_BINOP(div, "divss")
#undef RZ_BINOP

static inline float sqrt(float a) {
    uint32_t saved, mode;
    float r;
    __asm__ volatile("stmxcsr %0" : "=m"(saved));
    mode = saved | 0x6000u;
    __asm__ volatile("ldmxcsr %2\n\tsqrtss %1, %0\n\tldmxcsr %3" : "=x"(r) : "x"(a), "m"(mode), "m"(saved));
    return r;
}

// __fmaf_rz: the product is exact in double, the double add and the narrowing both truncate,
// and truncation composes, so this is the correctly truncated a*b+c.
static inline float fma(float a, float b, float c) {
    uint32_t saved, mode;
    double p = double(a) * double(b), s = double(c);
    float r;
    __asm__ volatile("stmxcsr %0" : "=m"(saved));
    mode = saved | 0x6000u;
    __asm__ volatile("ldmxcsr %3\n\taddsd %2, %1\n\tcvtsd2ss %1, %0\n\tldmxcsr %4"
                     : "=x"(r), "+x"(p)
                     : "x"(s), "m"(mode), "m"(saved));
    return r;
}

#endif // ORACLE_RZ_AS_RN

// __saturatef: clamp to [0, 1], NaN -> 0
static inline float saturate(float x) { return x > 0.f ? (x < 1.f ? x : 1.f) : 0.f; }

} // namespace rz
