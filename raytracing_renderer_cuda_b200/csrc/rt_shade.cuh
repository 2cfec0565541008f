// rt_shade.cuh — camera rays, textures (incl. Perlin), materials and the per-hit
// integrator step.  Semantics follow SURVEY.md §8a / §8a' item by item; every function
// cites the reference lines it restates.
#pragma once

#include "rt_intersect.cuh"

namespace rtd {

// ------------------------------------------------------------------ Perlin ----
// perlin_noise() is one out-of-line function and the octave loop is not unrolled: the shade kernel drops from
// 9 000 to 3 900 SASS instructions (144 KB -> 62 KB), which removes the instruction-cache stalls ncu showed
// (profiles/r01_wavefront_ncu.md) and lets ptxas fit 4 CTAs of 256 threads per SM (64 registers).
#ifndef RT_PERLIN_INLINE
#define RT_PERLIN_FN static __device__ __noinline__
#define RT_PERLIN_UNROLL _Pragma("unroll 1")
#else
#define RT_PERLIN_FN RT_DEV
#define RT_PERLIN_UNROLL _Pragma("unroll")
#endif
// Ken Perlin's 2002 permutation (perlin_noise.h:24-37).  p[512] of the reference is
// this table twice (perlin_noise.h:43), so p[i] == perm[i & 255] for every index used.
__device__ const uint8_t k_perlin_perm[256] = {
    151, 160, 137, 91,  90,  15,  131, 13,  201, 95,  96,  53,  194, 233, 7,   225, 140, 36,  103, 30,  69,  142,
    8,   99,  37,  240, 21,  10,  23,  190, 6,   148, 247, 120, 234, 75,  0,   26,  197, 62,  94,  252, 219, 203,
    117, 35,  11,  32,  57,  177, 33,  88,  237, 149, 56,  87,  174, 20,  125, 136, 171, 168, 68,  175, 74,  165,
    71,  134, 139, 48,  27,  166, 77,  146, 158, 231, 83,  111, 229, 122, 60,  211, 133, 230, 220, 105, 92,  41,
    55,  46,  245, 40,  244, 102, 143, 54,  65,  25,  63,  161, 1,   216, 80,  73,  209, 76,  132, 187, 208, 89,
    18,  169, 200, 196, 135, 130, 116, 188, 159, 86,  164, 100, 109, 198, 173, 186, 3,   64,  52,  217, 226, 250,
    124, 123, 5,   202, 38,  147, 118, 126, 255, 82,  85,  212, 207, 206, 59,  227, 47,  16,  58,  17,  182, 189,
    28,  42,  223, 183, 170, 213, 119, 248, 152, 2,   44,  154, 163, 70,  221, 153, 101, 155, 167, 43,  172, 9,
    129, 22,  39,  253, 19,  98,  108, 110, 79,  113, 224, 232, 178, 185, 112, 104, 218, 246, 97,  228, 251, 34,
    242, 193, 238, 210, 144, 12,  191, 179, 162, 241, 81,  51,  145, 235, 249, 14,  239, 107, 49,  192, 214, 31,
    181, 199, 106, 157, 184, 84,  204, 176, 115, 121, 50,  45,  127, 4,   150, 254, 138, 236, 205, 93,  222, 114,
    67,  29,  24,  72,  243, 141, 128, 195, 78,  66,  215, 61,  156, 180};
#ifdef RT_PERLIN_GLOBAL
// A/B: the pair table (perm[i] | perm[i + 1] << 8) read straight from global memory through L1 (1 KB), no staging
__device__ const uint32_t k_perlin_pairs[256] = {
    0xa097, 0x89a0, 0x5b89, 0x5a5b, 0x0f5a, 0x830f, 0x0d83, 0xc90d, 0x5fc9, 0x605f, 0x3560, 0xc235, 0xe9c2, 0x07e9, 0xe107, 0x8ce1,
    0x248c, 0x6724, 0x1e67, 0x451e, 0x8e45, 0x088e, 0x6308, 0x2563, 0xf025, 0x15f0, 0x0a15, 0x170a, 0xbe17, 0x06be, 0x9406, 0xf794,
    0x78f7, 0xea78, 0x4bea, 0x004b, 0x1a00, 0xc51a, 0x3ec5, 0x5e3e, 0xfc5e, 0xdbfc, 0xcbdb, 0x75cb, 0x2375, 0x0b23, 0x200b, 0x3920,
    0xb139, 0x21b1, 0x5821, 0xed58, 0x95ed, 0x3895, 0x5738, 0xae57, 0x14ae, 0x7d14, 0x887d, 0xab88, 0xa8ab, 0x44a8, 0xaf44, 0x4aaf,
    0xa54a, 0x47a5, 0x8647, 0x8b86, 0x308b, 0x1b30, 0xa61b, 0x4da6, 0x924d, 0x9e92, 0xe79e, 0x53e7, 0x6f53, 0xe56f, 0x7ae5, 0x3c7a,
    0xd33c, 0x85d3, 0xe685, 0xdce6, 0x69dc, 0x5c69, 0x295c, 0x3729, 0x2e37, 0xf52e, 0x28f5, 0xf428, 0x66f4, 0x8f66, 0x368f, 0x4136,
    0x1941, 0x3f19, 0xa13f, 0x01a1, 0xd801, 0x50d8, 0x4950, 0xd149, 0x4cd1, 0x844c, 0xbb84, 0xd0bb, 0x59d0, 0x1259, 0xa912, 0xc8a9,
    0xc4c8, 0x87c4, 0x8287, 0x7482, 0xbc74, 0x9fbc, 0x569f, 0xa456, 0x64a4, 0x6d64, 0xc66d, 0xadc6, 0xbaad, 0x03ba, 0x4003, 0x3440,
    0xd934, 0xe2d9, 0xfae2, 0x7cfa, 0x7b7c, 0x057b, 0xca05, 0x26ca, 0x9326, 0x7693, 0x7e76, 0xff7e, 0x52ff, 0x5552, 0xd455, 0xcfd4,
    0xcecf, 0x3bce, 0xe33b, 0x2fe3, 0x102f, 0x3a10, 0x113a, 0xb611, 0xbdb6, 0x1cbd, 0x2a1c, 0xdf2a, 0xb7df, 0xaab7, 0xd5aa, 0x77d5,
    0xf877, 0x98f8, 0x0298, 0x2c02, 0x9a2c, 0xa39a, 0x46a3, 0xdd46, 0x99dd, 0x6599, 0x9b65, 0xa79b, 0x2ba7, 0xac2b, 0x09ac, 0x8109,
    0x1681, 0x2716, 0xfd27, 0x13fd, 0x6213, 0x6c62, 0x6e6c, 0x4f6e, 0x714f, 0xe071, 0xe8e0, 0xb2e8, 0xb9b2, 0x70b9, 0x6870, 0xda68,
    0xf6da, 0x61f6, 0xe461, 0xfbe4, 0x22fb, 0xf222, 0xc1f2, 0xeec1, 0xd2ee, 0x90d2, 0x0c90, 0xbf0c, 0xb3bf, 0xa2b3, 0xf1a2, 0x51f1,
    0x3351, 0x9133, 0xeb91, 0xf9eb, 0x0ef9, 0xef0e, 0x6bef, 0x316b, 0xc031, 0xd6c0, 0x1fd6, 0xb51f, 0xc7b5, 0x6ac7, 0x9d6a, 0xb89d,
    0x54b8, 0xcc54, 0xb0cc, 0x73b0, 0x7973, 0x3279, 0x2d32, 0x7f2d, 0x047f, 0x9604, 0xfe96, 0x8afe, 0xec8a, 0xcdec, 0x5dcd, 0xde5d,
    0x72de, 0x4372, 0x1d43, 0x181d, 0x4818, 0xf348, 0x8df3, 0x808d, 0xc380, 0x4ec3, 0x424e, 0xd742, 0x3dd7, 0x9c3d, 0xb49c, 0x97b4};
#endif


// Shared-memory staging of the permutation.  Entry i holds the PAIR (perm[i], perm[i+1])
// so the two neighbouring lookups every hash level needs (perlin_noise.h:67-72) cost one
// load; the table is replicated 2^RT_PERLIN_BANK_BITS times (word i * replicas + lane mod replicas).
// Optionally (RT_PERLIN_GTAB) a second table holds the 16 gradient directions of perlin_noise::grad as float4 coefficient vectors (one
// LDS.128 per corner, replicated per lane: 16 x 32 x 16 B = 8 KB), see perlin_grad.
#ifndef RT_PERLIN_BANK_BITS
// Replicas per entry = 2^bits (lane l reads replica l mod 2^bits).  One replica per bank (5 bits, 32 KB per CTA) makes
// every lookup conflict-free, but the table is staged by every CTA of every launch and its shared memory comes out of
// the L1 that holds the kernels' stack frames: measured on C1 / C3 (gpurun_out/ab_banks.log, ab_banks2.log, 128-thread
// CTAs): 5 bits 8.92 / 4.79 ms, 4: 8.75, 3: 8.69, 2: 8.65, 1: 8.58 / 4.69, 0: 8.68 / 4.70; straight from global memory
// through L1 (RT_PERLIN_GLOBAL): 8.52 / 4.77.  Two replicas (2 KB) it is.
#define RT_PERLIN_BANK_BITS 1
#endif
#define RT_PERLIN_PERM_WORDS (256 << RT_PERLIN_BANK_BITS)
#ifdef RT_PERLIN_GTAB // opt-in: measured 1 % SLOWER on C1 than the compare/select form (gpurun_out/ab_perlin.log)
#define RT_PERLIN_SMEM_WORDS (RT_PERLIN_PERM_WORDS + 16 * 32 * 4)
#else
#define RT_PERLIN_SMEM_WORDS RT_PERLIN_PERM_WORDS
#endif
struct PerlinTab {
    const uint32_t* s; // shared memory, RT_PERLIN_SMEM_WORDS words
    uint32_t lane;
};
RT_DEV void perlin_stage(uint32_t* smem, uint32_t tid, uint32_t nthreads) {
#ifdef RT_PERLIN_GLOBAL
    (void)smem; (void)tid; (void)nthreads;
    return;
#endif
#if RT_PERLIN_BANK_BITS >= 2
    // 4 consecutive replica words hold the same pair: one 16-byte store per 4 words
    uint4* s4 = reinterpret_cast<uint4*>(smem);
    for (uint32_t w = tid; w < RT_PERLIN_PERM_WORDS / 4; w += nthreads) {
        uint32_t i = w >> (RT_PERLIN_BANK_BITS - 2);
        uint32_t v = uint32_t(k_perlin_perm[i]) | (uint32_t(k_perlin_perm[(i + 1) & 255]) << 8);
        s4[w] = make_uint4(v, v, v, v);
    }
#else
    for (uint32_t w = tid; w < RT_PERLIN_PERM_WORDS; w += nthreads) {
        uint32_t i = w >> RT_PERLIN_BANK_BITS;
        smem[w] = uint32_t(k_perlin_perm[i]) | (uint32_t(k_perlin_perm[(i + 1) & 255]) << 8);
    }
#endif
#ifdef RT_PERLIN_GTAB
    // gradient h = hash & 15 (perlin_noise.h:173-181): u = h<8 ? x : y; v = h<4 ? y : (h==12||h==14 ? x : z);
    // grad = (h&1 ? -u : u) + (h&2 ? -v : v)  ==  cx*x + cy*y + cz*z with two coefficients +-1 and one 0
    float4* g4 = reinterpret_cast<float4*>(smem + RT_PERLIN_PERM_WORDS);
    for (uint32_t e = tid; e < 16u * 32u; e += nthreads) {
        const uint32_t h = e >> 5;
        float c[3] = {0.f, 0.f, 0.f};
        const int ua = h < 8u ? 0 : 1;
        const int va = h < 4u ? 1 : ((h == 12u || h == 14u) ? 0 : 2);
        c[ua] = (h & 1u) ? -1.f : 1.f;
        c[va] = (h & 2u) ? -1.f : 1.f;
        g4[e] = make_float4(c[0], c[1], c[2], 0.f);
    }
#endif
}
RT_DEV uint32_t perlin_pair(const PerlinTab& pt, uint32_t i) {
#ifdef RT_PERLIN_GLOBAL
    return __ldg(&k_perlin_pairs[i & 255u]);
#endif
    return pt.s[((i & 255u) << RT_PERLIN_BANK_BITS) | (pt.lane & ((1u << RT_PERLIN_BANK_BITS) - 1u))];
}

// perlin_noise::grad (perlin_noise.h:173-181)
#ifdef RT_PERLIN_GTAB
// Table form: exact — the zero coefficient adds nothing and the two unit coefficients give the same single rounding
// as +-u +-v — and 3 FP instructions + one LDS.128 instead of ~13 compare/select/logic instructions per corner.
RT_DEV float perlin_grad(const PerlinTab& pt, uint32_t hash, float x, float y, float z) {
    const float4 c = reinterpret_cast<const float4*>(pt.s + RT_PERLIN_PERM_WORDS)[((hash & 15u) << 5) | pt.lane];
    return __fmaf_rn(c.z, z, __fmaf_rn(c.y, y, c.x * x));
}
#else
RT_DEV float perlin_grad(const PerlinTab&, uint32_t hash, float x, float y, float z) {
    uint32_t h = hash & 15u;
    float u = h < 8u ? x : y;
    float v = h < 4u ? y : ((h == 12u || h == 14u) ? x : z);
    return ((h & 1u) == 0u ? u : -u) + ((h & 2u) == 0u ? v : -v);
}
#endif
RT_DEV float perlin_ease(float t) { return t * t * t * (t * (t * 6.f - 15.f) + 10.f); } // :156-165
RT_DEV float perlin_lerp(float t, float a, float b) { return a + t * (b - a); }         // :167-171

// perlin_noise::noise (perlin_noise.h:46-105)
RT_DEV float perlin_noise_body(const PerlinTab& pt, V3 p) {
    float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    uint32_t xi = uint32_t(int(fx)) & 255u, yi = uint32_t(int(fy)) & 255u, zi = uint32_t(int(fz)) & 255u;
    float xf = p.x - fx, yf = p.y - fy, zf = p.z - fz;
    float u = perlin_ease(xf), v = perlin_ease(yf), w = perlin_ease(zf);
    uint32_t px = perlin_pair(pt, xi);        // (p[xi], p[xi+1])
    uint32_t A = (px & 255u) + yi;            // p[xi] + yi
    uint32_t B = ((px >> 8) & 255u) + yi;     // p[xi+1] + yi
    uint32_t pa = perlin_pair(pt, A);         // (p[A], p[A+1])
    uint32_t pb = perlin_pair(pt, B);         // (p[B], p[B+1])
    uint32_t AA = (pa & 255u) + zi, AB = ((pa >> 8) & 255u) + zi;
    uint32_t BA = (pb & 255u) + zi, BB = ((pb >> 8) & 255u) + zi;
    uint32_t gaa = perlin_pair(pt, AA), gba = perlin_pair(pt, BA); // (p[AA], p[AA+1]) ...
    uint32_t gab = perlin_pair(pt, AB), gbb = perlin_pair(pt, BB);
    float x1 = xf - 1.f, y1 = yf - 1.f, z1 = zf - 1.f;
    float res = perlin_lerp(
        w,
        perlin_lerp(v, perlin_lerp(u, perlin_grad(pt, gaa, xf, yf, zf), perlin_grad(pt, gba, x1, yf, zf)),
                    perlin_lerp(u, perlin_grad(pt, gab, xf, y1, zf), perlin_grad(pt, gbb, x1, y1, zf))),
        perlin_lerp(v, perlin_lerp(u, perlin_grad(pt, gaa >> 8, xf, yf, z1), perlin_grad(pt, gba >> 8, x1, yf, z1)),
                    perlin_lerp(u, perlin_grad(pt, gab >> 8, xf, y1, z1), perlin_grad(pt, gbb >> 8, x1, y1, z1))));
    return (res + 1.0f) / 2.0f;
}
RT_PERLIN_FN float perlin_noise(const PerlinTab& pt, V3 p) { return perlin_noise_body(pt, p); }

// perlin_noise::turbulance_noise, implementation 3 (perlin_noise.h:142-153), defaults
// lacunacity 2, gain .5, 6 octaves (perlin_noise.h:13-17)
#ifdef RT_PERLIN_PAIR
// two octaves per call: the two lattice walks (floor -> three dependent table levels -> eight gradients -> seven
// lerps) are independent, so their fixed-latency chains interleave; the sum is still taken octave by octave
struct F2 {
    float a, b;
};
RT_PERLIN_FN F2 perlin_noise2(const PerlinTab& pt, V3 p, float f0, float f1) {
    return F2{perlin_noise_body(pt, p * f0), perlin_noise_body(pt, p * f1)};
}
RT_DEV float perlin_turbulence(const PerlinTab& pt, V3 p) {
    float frequency = 1.f, sum = 0.f, amplitude = 1.f;
    RT_PERLIN_UNROLL
    for (int i = 0; i < 3; ++i) {
        const F2 r = perlin_noise2(pt, p, frequency, frequency * 2.f);
        sum += fabsf(r.a * 2.f - 1.f) * amplitude;
        sum += fabsf(r.b * 2.f - 1.f) * (amplitude * 0.5f);
        frequency *= 4.f;
        amplitude *= 0.25f;
    }
    return sum;
}
#else
#if !defined(RT_PERLIN_INLINE) && !defined(RT_PERLIN_TURB_CALLS)
// ONE out-of-line call per 6-octave evaluation with the noise body inlined in its loop (six calls of perlin_noise cost
// 1-2 % more on C1/C3: gpurun_out/ab_turbfn.log)
static __device__ __noinline__ float perlin_turbulence(const PerlinTab& pt, V3 p) {
    float frequency = 1.f, sum = 0.f, amplitude = 1.f;
#pragma unroll 1
    for (int i = 0; i < 6; ++i) {
        float r = perlin_noise_body(pt, p * frequency);
        sum += fabsf(r * 2.f - 1.f) * amplitude;
        frequency *= 2.f;
        amplitude *= 0.5f;
    }
    return sum;
}
#else
RT_DEV float perlin_turbulence(const PerlinTab& pt, V3 p) {
    float frequency = 1.f, sum = 0.f, amplitude = 1.f;
    RT_PERLIN_UNROLL
    for (int i = 0; i < 6; ++i) {
        float r = perlin_noise(pt, p * frequency);
        sum += fabsf(r * 2.f - 1.f) * amplitude;
        frequency *= 2.f;
        amplitude *= 0.5f;
    }
    return sum;
}
#endif
#endif

// ------------------------------------------------------------------ textures ----
RT_DEV DTexture load_tex(const DScene& sc, int32_t ix) {
    const float4* tp = reinterpret_cast<const float4*>(sc.texs + ix);
    float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
    DTexture t;
    t.kind = __float_as_uint(a.x);
    t.even = __float_as_int(a.y);
    t.odd = __float_as_int(a.z);
    t.image = __float_as_int(a.w);
    t.c1x = b.x; t.c1y = b.y; t.c1z = b.z; t.density = b.w;
    t.c2x = c.x; t.c2y = c.y; t.c2z = c.z; t.hardness = c.w;
    return t;
}

// checker_texture::value (texture.h:41-48) only selects a child from p; children may be
// checkers themselves.  Returns the index of the leaf texture and loads it into `t`.
RT_DEV int32_t resolve_texture(const DScene& sc, int32_t ix, V3 p, DTexture& t) {
    t = load_tex(sc, ix);
    for (int guard = 0; guard < 8 && t.kind == RT_TEX_CHECKER; ++guard) {
        float sines = __sinf(10 * p.x) * __sinf(10 * p.y) * __sinf(10 * p.z);
        ix = sines < 0.f ? t.odd : t.even;
        t = load_tex(sc, ix);
    }
    return ix;
}

RT_DEV V3 tex_constant(const DTexture& t) { return mk(t.c1x, t.c1y, t.c1z); } // texture.h:22-24
RT_DEV V3 tex_perlin(const PerlinTab& pt, const DTexture& t, V3 p) {         // texture.h:58-59
    float v = perlin_noise(pt, p * t.density);
    return mk(v, v, v);
}
RT_DEV V3 tex_turbulence(const PerlinTab& pt, const DTexture& t, V3 p) { // texture.h:60-63: vec3(1) * 0.5 * turb(p * density)
    float v = __fmul_rz(0.5f, perlin_turbulence(pt, p * t.density));
    return mk(v, v, v);
}
RT_DEV V3 tex_marble(const PerlinTab& pt, const DTexture& t, V3 p) { // texture.h:65-75: turbulence of the UNSCALED p
    float value = 0.5f * (1 + __sinf((p.z * t.density + 7 * perlin_turbulence(pt, p))));
    V3 color1 = mk(0.925f, 0.816f, 0.78f);
    V3 color2 = mk(float(0.349 / 2), float(0.431 / 2), float(0.498 / 2));
    return color1 * value + color2 * (1 - value);
}
RT_DEV V3 tex_wood(const PerlinTab& pt, const DTexture& t, V3 p) { // texture.h:99-104
    float nn = t.hardness * perlin_noise(pt, p / t.density);
    nn -= floorf(nn);
    return (mk(t.c1x, t.c1y, t.c1z) * nn) + (mk(t.c2x, t.c2y, t.c2z) * (1.f - nn));
}
RT_DEV V3 tex_image(const DScene& sc, const DTexture& t, V3 n) { // image_texture::value (texture.h:118-132)
    float u, v;
    sphere_uv(n, u, v);
    const DImage& im = sc.images[t.image];
    int i = u * im.width;
    int j = (1 - v) * im.height - 0.001;
    if (i < 0) i = 0;
    if (j < 0) j = 0;
    if (i > im.width - 1) i = im.width - 1;
    if (j > im.height - 1) j = im.height - 1;
    float4 c = tex2D<float4>(im.tex, i + 0.5f, j + 0.5f); // nearest texel, row 0 = top
    return mk(c.x, c.y, c.z);
}

// text::value(u, v, p) of a non-checker texture.  (u,v) are only needed by image textures
// and are derived from the outward normal there (sphere.h:123-128).
RT_DEV V3 texture_leaf_value(const DScene& sc, const PerlinTab& pt, const DTexture& t, V3 n, V3 p) {
    switch (t.kind) {
    case RT_TEX_CONSTANT: return tex_constant(t);
    case RT_TEX_NOISE_PERLIN: return tex_perlin(pt, t, p);
    case RT_TEX_NOISE_TURBULANCE: return tex_turbulence(pt, t, p);
    case RT_TEX_NOISE_MARBLE: return tex_marble(pt, t, p);
    case RT_TEX_WOOD: return tex_wood(pt, t, p);
    case RT_TEX_IMAGE: return tex_image(sc, t, n);
    default: return mk(1.f, 1.f, 1.f);
    }
}

RT_DEV V3 texture_value(const DScene& sc, const PerlinTab& pt, int32_t ix, V3 n, V3 p) {
    DTexture t;
    resolve_texture(sc, ix, p, t);
    return texture_leaf_value(sc, pt, t, n, p);
}

// ------------------------------------------------------------------ camera ----
// render()'s per-sample set-up (main.cu:116-118) + camera::get_ray (camera.h:33-38).
// Draw order of the reference: jitter x, jitter y, lens disk, shutter time.
RT_DEV Ray camera_ray(const DScene& sc, const DRenderParams& rp, uint32_t pixel, uint32_t sample) {
    const DCamera& cam = sc.cam;
    uint32_t i = pixel % uint32_t(rp.width), j = pixel / uint32_t(rp.width);
    U4 r0 = rng_block(rp.seed, pixel, sample, 0, 0);
    float s = float(i + u01(r0.x)) / float(rp.width);
    float t = float(j + u01(r0.y)) / float(rp.height);
    float dx, dy;
    sample_unit_disk(u01(r0.z), u01(r0.w), dx, dy);
    V3 rd = cam.lens_radius * mk(dx, dy, 0.f);
    V3 offset = cam.u * rd.x + cam.v * rd.y;
    float time = cam.t0;
    if (cam.t1 != cam.t0) {
        U4 r1 = rng_block(rp.seed, pixel, sample, 0, 1);
        time = __fadd_rn(cam.t0, __fmul_rn(u01(r1.x), __fsub_rn(cam.t1, cam.t0))); // single roundings: the same shutter time in every kernel
    }
    Ray r;
    r.o = cam.origin + offset;
    r.d = cam.lower_left + s * cam.horizontal + t * cam.vertical - cam.origin - offset;
    r.time = time;
    return r;
}

// ------------------------------------------------------------------ materials ----
RT_DEV DMaterial load_mat(const DScene& sc, uint32_t ix) {
    const float4* mp = reinterpret_cast<const float4*>(sc.mats + ix);
    float4 a = __ldg(mp), b = __ldg(mp + 1);
    DMaterial m;
    m.kind = __float_as_uint(a.x);
    m.tex = __float_as_int(a.y);
    m.ax = a.z; m.ay = a.w; m.az = b.x; m.param = b.y;
    m.pad0 = m.pad1 = 0.f;
    return m;
}

RT_DEV V3 reflect(V3 v, V3 n) { return v - 2.f * dot(v, n) * n; } // utils.h:93-97

// utils::refract (utils.h:107-122)
RT_DEV bool refract(V3 v, V3 n, float mu, V3& refracted) {
    V3 i = normalize(v);
    float in = dot(i, n);
    // Written with explicit single roundings: left to the compiler, `1 - in * in` was fused into an FMA in some kernels
    // and not in others (same source, different inlining context), and the refracted rays of the persistent-lane kernel
    // left the megakernel's by an ulp — 3.8 % of the pixels of the million-sphere frame (gpurun_out/pt_debug.log).
    float delta = __fsub_rn(1.f, __fmul_rn(__fmul_rn(mu, mu), __fsub_rn(1.f, __fmul_rn(in, in))));
    if (delta > 0) {
        refracted = mu * (i - n * in) - n * sqrtf(delta);
        return true;
    }
    return false;
}

// utils::shlick (utils.h:124-137)
RT_DEV float shlick(float cosine, float ref_id) {
    float r0 = __fdiv_rz((1.f - ref_id), (1.f + ref_id));
    r0 = __fmul_rz(r0, r0);
    return r0 + __fmul_rz((1.f - r0), __powf(1.f - cosine, 5.f));
}

// lambertian::scatter (material.h:105-116): target = p + n + ball; ray keeps rin's time
RT_DEV void scatter_lambertian(const RayQ& q, V3 p, V3 n, U4 r, Ray& out) {
    V3 target = p + n + sample_unit_ball(u01(r.x), u01(r.y), u01(r.z));
    out.o = p;
    out.d = target - p;
    out.time = q.time;
}

// ---- emitter importance sampling (RT_RENDER_EMITTER_SAMPLING; the reference README's roadmap item, README.md:27-28) ----
// In color() (main.cu:39-57) a path that reaches an emitter is worth emit + bloom, whatever it did before.  At a
// lambertian hit the reference draws d = n + uniform-in-ball (material.h:112), whose direction density about the
// normal is p_ref(w) = 2 cos^3(theta) / pi (the chord of the unit ball centred at n along w is 2 cos(theta); |d|
// along it has a density proportional to t^2).  The value of the vertex splits into
//     V = Int p_ref [next hit is a listed emitter] (emit + bloom)  +  Int p_ref [anything else] (rest of color())
// With the flag on, the first integral is estimated by ONE extra "shadow/emission" ray per lambertian hit, aimed
// uniformly into the cone of an emitter sphere chosen by solid angle (|d| from the reference's conditional law given w,
// 2 cos(theta) cbrt(u): tmin is in units of |d|), worth (p_ref / p_sel) (emit + bloom) if its closest hit is an
// emitter; the second by the reference's own scattered ray, which now counts 0 when its next hit is a listed
// emitter.  Same expectation, weights p_ref / p_sel of the order of an emitter's solid angle — never above the
// reference estimator's own range.  Ordinary colour-math FP32; nothing here is compared bit for bit.
struct LightCone {
    V3 axis;             // unit vector from p to the emitter's centre
    float one_minus_cos; // w is inside the cone  <=>  1 - dot(w, axis) <= one_minus_cos   (2 = every direction)
};
RT_DEV LightCone light_cone(const DScene& sc, uint32_t k, V3 p, float time) {
    const uint32_t prim = sc.lights[k];
    const float4 a = __ldg(&sc.sph_a[prim]);
    V3 c = mk(a.x, a.y, a.z);
    if (prim >= sc.n_static) c = moving_center(a, __ldg(&sc.sph_b[prim]), __uint_as_float(__ldg(&sc.sph_c[prim]).x), time);
    const float lx = c.x - p.x, ly = c.y - p.y, lz = c.z - p.z;
    const float d2 = lx * lx + ly * ly + lz * lz, r2 = a.w * a.w;
    LightCone lc;
    if (!(d2 > r2)) { // on or inside the emitter: every direction reaches it
        lc.axis = mk(0.f, 0.f, 1.f);
        lc.one_minus_cos = 2.f;
    } else {
        const float inv = 1.f / sqrtf(d2);
        lc.axis = V3{lx * inv, ly * inv, lz * inv};
        const float s2 = r2 / d2; // sin^2 of the half angle
        lc.one_minus_cos = fmaxf(s2 / (1.f + sqrtf(1.f - s2)), 1e-12f);
    }
    return lc;
}
// Is `prim` one of the emitter spheres the shadow rays sample?  (All of them unless the scene has more than RT_MAX_LIGHTS.)
RT_DEV bool light_listed(const DScene& sc, uint32_t prim) {
    bool listed = false;
    for (uint32_t k = 0; k < sc.n_lights; ++k) listed = listed || sc.lights[k] == prim;
    return listed;
}
// The shadow ray's direction `d` (with the reference's |d| law) at a lambertian hit (p, n); `rl` = Philox block (bounce, 2).
// Emitter k is chosen with probability proportional to its solid angle, the direction uniformly inside its cone, so
// p_sel(w) = (number of cones holding w) / (total solid angle) and p_ref / p_sel <= p_ref x total solid angle.
// Returns p_ref / p_sel, or 0 when the direction leaves below the horizon (p_ref = 0 there: no ray is traced).
RT_DEV float sample_light_direction(const DScene& sc, V3 p, V3 n, float time, U4 rl, V3& d) {
    const float inv_n = 1.f / sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
    float total = 0.f; // sum of (1 - cos half-angle) = total solid angle / 2 pi
    for (uint32_t j = 0; j < sc.n_lights; ++j) total += light_cone(sc, j, p, time).one_minus_cos;
    const float pick = u01(rl.x) * total;
    LightCone lc = light_cone(sc, 0u, p, time);
    uint32_t k = 0;
    float below = 0.f;
    for (uint32_t j = 0; j + 1u < sc.n_lights && !(pick <= below + lc.one_minus_cos); ++j) { // walk the cumulative sum
        below += lc.one_minus_cos;
        k = j + 1u;
        lc = light_cone(sc, k, p, time);
    }
    const float cz = 1.f - u01(rl.y) * lc.one_minus_cos; // cos of the angle to the axis, uniform over the cap
    const float sz = sqrtf(fmaxf(0.f, 1.f - cz * cz));
    float sn, cs;
    __sincosf(6.283185307179586f * u01(rl.z), &sn, &cs);
    const V3 ax = lc.axis; // orthonormal basis around it (Duff et al. 2017)
    const float sg = ax.z >= 0.f ? 1.f : -1.f;
    const float ka = -1.f / (sg + ax.z), kb = ax.x * ax.y * ka;
    const V3 t1 = V3{1.f + sg * ax.x * ax.x * ka, sg * kb, -sg * ax.x};
    const V3 t2 = V3{kb, sg + ax.y * ax.y * ka, -ax.y};
    const float e1 = sz * cs, e2 = sz * sn;
    const V3 w = V3{e1 * t1.x + e2 * t2.x + cz * ax.x, e1 * t1.y + e2 * t2.y + cz * ax.y, e1 * t1.z + e2 * t2.z + cz * ax.z};
    const float cos_t = (w.x * n.x + w.y * n.y + w.z * n.z) * inv_n;
    if (!(cos_t > 0.f)) return 0.f;
    const float len = 2.f * cos_t * fminf(cbrtf(u01(rl.w)), 0.99999994f);
    d = V3{w.x * len, w.y * len, w.z * len};
    const float p_ref = 0.6366197723675814f * cos_t * cos_t * cos_t; // 2 cos^3 / pi
    float holding = 0.f; // cones that hold w (the chosen one does by construction)
    for (uint32_t j = 0; j < sc.n_lights; ++j) {
        const LightCone lj = light_cone(sc, j, p, time);
        const float off = 1.f - (w.x * lj.axis.x + w.y * lj.axis.y + w.z * lj.axis.z);
        if (j == k || lj.one_minus_cos >= 2.f || off <= lj.one_minus_cos) holding += 1.f;
    }
    return p_ref * 6.283185307179586f * total / holding;
}
// The shadow/emission ray of one lambertian hit: (p_ref / p_sel) (emit + bloom) if its closest hit is a listed emitter.
RT_DEV V3 emitter_sample(const DScene& sc, const DRenderParams& rp, const PerlinTab& pt, V3 p, V3 n, float time, uint32_t pixel,
                         uint32_t sample, uint32_t bounce, uint32_t& rays) {
    V3 d;
    const float w = sample_light_direction(sc, p, n, time, rng_block(rp.seed, pixel, sample, bounce, 2), d);
    if (w == 0.f) return mk(0.f, 0.f, 0.f);
    Ray r;
    r.o = p;
    r.d = d;
    r.time = time;
    const RayQ q = make_rayq(r);
    const Hit h = closest_hit(sc, q, rp.tmin, true);
    ++rays;
    if (h.prim == RT_INVALID_ID) return mk(0.f, 0.f, 0.f);
    const DMaterial m = load_mat(sc, __ldg(&sc.sph_c[h.prim]).y);
    if (m.kind != RT_MAT_EMITTER || !light_listed(sc, h.prim)) return mk(0.f, 0.f, 0.f);
    V3 hp, hn;
    hit_surface(sc, q, h, hp, hn);
    const V3 e = texture_value(sc, pt, m.tex, hn, hp) * m.param + mk(rp.bloom, rp.bloom, rp.bloom); // emitter::emit + bloom
    return V3{e.x * w, e.y * w, e.z * w};
}
RT_DEV bool is_listed_emitter(const DScene& sc, uint32_t prim) {
    const float4 m0 = __ldg(reinterpret_cast<const float4*>(sc.mats + __ldg(&sc.sph_c[prim]).y));
    return __float_as_uint(m0.x) == RT_MAT_EMITTER && light_listed(sc, prim);
}

// metal::scatter (material.h:118-131): draws the ball sample even at roughness 0, ray time
// resets to 0 (ray.h:12 default), absorbed when dot(out, n) <= 0
RT_DEV bool scatter_metal(const RayQ& q, V3 p, V3 n, float roughness, U4 r, Ray& out) {
    V3 reflection = reflect(normalize(q.d), n);
    out.o = p;
    out.d = reflection + roughness * sample_unit_ball(u01(r.x), u01(r.y), u01(r.z));
    out.time = 0.f;
    return dot(out.d, n) > 0.f;
}

// dielectric::scatter (material.h:133-184); ray time resets to 0
RT_DEV void scatter_dielectric(const RayQ& q, V3 p, V3 n, float ri, U4 r, Ray& out) {
    V3 refraction_normal;
    V3 reflected = reflect(q.d, n);
    float mu, cosine;
    float ddn = dot(q.d, n);
    if (ddn > 0.f) {
        refraction_normal = -n;
        mu = ri;
        cosine = __fdiv_rn(ddn, length(q.d));
        cosine = __fsqrt_rz(__fsub_rn(1.f, __fmul_rn(__fmul_rn(ri, ri), __fsub_rn(1.f, __fmul_rn(cosine, cosine))))); // no contraction, as above
    } else {
        refraction_normal = n;
        mu = 1.f / ri;
        cosine = __fdiv_rn(-ddn, length(q.d));
    }
    float reflect_prob;
    V3 refracted = mk(0.f, 0.f, 0.f);
    if (refract(q.d, refraction_normal, mu, refracted)) {
        reflect_prob = shlick(cosine, ri);
    } else {
        reflect_prob = 1.f;
    }
    out.o = p;
    out.d = (u01(r.w) < reflect_prob) ? reflected : refracted;
    out.time = 0.f;
}

// The terms of one integrator step at an accepted hit — the body of color()'s loop (main.cu:45-55):
//   E = m.emit(h) + bloom;  `att`/`out` = what m.scatter(...) produces.  Returns false when scatter() does
// (emitter, absorbed metal ray): the path's value is then E.  `bounce` counts from 1 for the RNG key.
// RT_RENDER_EMITTER_SAMPLING: a lambertian hit also traces its shadow/emission ray — `direct` receives that ray's
// contribution, `rays` counts it, and `nee_vertex` tells the caller that a listed emitter hit by `out` is worth 0.
RT_DEV bool shade_terms(const DScene& sc, const DRenderParams& rp, const PerlinTab& pt, const RayQ& q, Hit h,
                        uint32_t pixel, uint32_t sample, uint32_t bounce, V3& E, V3& att, Ray& out, V3& direct, bool& nee_vertex,
                        uint32_t& rays) {
    V3 p, n;
    hit_surface(sc, q, h, p, n);
    DMaterial m = load_mat(sc, __ldg(&sc.sph_c[h.prim]).y);
    V3 bloom = mk(rp.bloom, rp.bloom, rp.bloom);
    att = mk(0.f, 0.f, 0.f);
    nee_vertex = false;
    if (m.kind == RT_MAT_EMITTER) { // emitter::emit / scatter (material.h:42-52)
        E = texture_value(sc, pt, m.tex, n, p) * m.param + bloom;
        return false;
    }
    E = mk(0.f, 0.f, 0.f) + bloom; // material::emit (material.h:14-16)
    U4 r = rng_block(rp.seed, pixel, sample, bounce, 0);
    if (m.kind == RT_MAT_LAMBERTIAN) {
        scatter_lambertian(q, p, n, r, out);
        att = texture_value(sc, pt, m.tex, n, p);
        // (the reference traces at most max_depth rays per path, main.cu:42: the shadow ray stands for trace bounce + 1)
        if ((rp.flags & RT_RENDER_EMITTER_SAMPLING) && sc.n_lights > 0u && int(bounce) < rp.max_depth) {
            const V3 c = emitter_sample(sc, rp, pt, p, n, q.time, pixel, sample, bounce, rays);
            direct = V3{direct.x + c.x, direct.y + c.y, direct.z + c.z};
            nee_vertex = true;
        }
        return true;
    }
    att = mk(m.ax, m.ay, m.az);
    if (m.kind == RT_MAT_METAL) return scatter_metal(q, p, n, m.param, r, out); // false: absorbed, the value is E
    scatter_dielectric(q, p, n, m.param, r, out);
    return true;
}

// A <- E + att*A (main.cu:51) or A <- E when the path ends.  Returns true if the path continues.
RT_DEV bool shade_hit(const DScene& sc, const DRenderParams& rp, const PerlinTab& pt, const RayQ& q, Hit h,
                      uint32_t pixel, uint32_t sample, uint32_t bounce, V3& A, Ray& out, V3& direct, bool& nee_vertex, uint32_t& rays) {
    V3 E, att;
    if (!shade_terms(sc, rp, pt, q, h, pixel, sample, bounce, E, att, out, direct, nee_vertex, rays)) {
        A = E;
        return false;
    }
    A = E + att * A;
    return true;
}

} // namespace rtd
