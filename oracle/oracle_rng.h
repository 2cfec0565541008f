// oracle/oracle_rng.h — TEST INFRASTRUCTURE.  The sequential generator the host builds use in
// place of cuRAND's XORWOW (main.cu:91,116-117; utils.h:69-72,84-85; camera.h:36;
// material.h:177): splitmix64, top 32 bits per draw.  Both the shim under the reference
// headers (oracle/shim/curand_kernel.h) and the restatement (oracle/rt_oracle.cpp, sampler 0)
// use it, so the two can be compared draw for draw.
#pragma once
#include <cstdint>

struct orng_state {
    uint64_t s;
};
static inline uint64_t orng_mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// consecutive seeds (the reference seeds pixel k with SEED + k, main.cu:91) must give unrelated
// streams, so the seed is hashed before it becomes the splitmix state
static inline void orng_init(unsigned long long seed, unsigned long long seq, unsigned long long offset, orng_state* st) {
    st->s = orng_mix(seed + 0x632BE59BD9B4E019ull) ^ orng_mix(seq * 0xD1B54A32D192ED03ull + offset + 1);
}
static inline uint32_t orng_next(orng_state* st) {
    uint64_t z = (st->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return uint32_t((z ^ (z >> 31)) >> 32);
}
// cuRAND's mapping of 32 bits to (0, 1]: x * 2^-32 + 2^-33 evaluated as one fused multiply-add
static inline float orng_uniform(orng_state* st) {
    return float(double(orng_next(st)) * 2.3283064365386963e-10 + 1.1641532182693481e-10);
}
