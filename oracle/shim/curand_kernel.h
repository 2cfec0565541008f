// oracle/shim/curand_kernel.h — TEST INFRASTRUCTURE (oracle O3).
// Lets the UNCHANGED reference headers compile for the host: common.h:9 includes this file by
// name, so putting this directory first on the include path supplies, in one place,
//   * empty __host__/__device__/__global__ qualifiers,
//   * the CUDA intrinsics the headers call, with exact round-toward-zero semantics (rz_math.h),
//   * __sinf/__tanf/__powf as libm calls (SFU approximations cannot be reproduced on a CPU),
//   * a curandState/curand_init/curand/curand_uniform built on oracle_rng.h.
// Compile with -D__CUDA_ARCH__=1000 so vec3.h:5-7 selects the intrinsic code path.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../oracle_rng.h"
#include "../rz_math.h"

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline

static inline float __fadd_rz(float a, float b) { return rz::add(a, b); }
static inline float __fsub_rz(float a, float b) { return rz::sub(a, b); }
static inline float __fmul_rz(float a, float b) { return rz::mul(a, b); }
static inline float __fdiv_rz(float a, float b) { return rz::div(a, b); }
static inline float __fsqrt_rz(float a) { return rz::sqrt(a); }
static inline float __fmaf_rz(float a, float b, float c) { return rz::fma(a, b, c); }
static inline float __saturatef(float a) { return rz::saturate(a); }
// glibc declares functions called __sinf/__tanf/__powf; macros sidestep the clash
#define __sinf(x) sinf(x)
#define __tanf(x) tanf(x)
#define __powf(x, y) powf(x, y)

typedef orng_state curandState;
static inline void curand_init(unsigned long long seed, unsigned long long seq, unsigned long long off, curandState* st) {
    orng_init(seed, seq, off, st);
}
static inline unsigned int curand(curandState* st) { return orng_next(st); }
static inline float curand_uniform(curandState* st) { return orng_uniform(st); }
