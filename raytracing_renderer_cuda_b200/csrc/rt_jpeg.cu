// rt_jpeg.cu — the reference's output stage on the device.
//
// The reference ends its frame on the host: a loop flips the rows and quantises the floats to bytes
// (main.cu:475-488), then the vendored stb_image_write v1.15 writes a baseline JPEG at quality 100
// (main.cu:491 -> stb_image_write.h:1368-1573, Jon Olick's jo_jpeg).  Here the bytes never leave the GPU
// until they are the finished file: k_tonemap produces the flipped rgb8 image, and the kernels below
// produce a JPEG stream that is BYTE-IDENTICAL to what stbi_write_jpg emits for the same pixels and
// quality (oracle: oracle/jpeg_oracle.cpp, pinned against the real stb through oracle/_ref).
//
// The sequential encoder (one MCU after the other through a 24-bit bit buffer, stb_image_write.h:1223-1238)
// becomes data-parallel passes:
//   k_jpeg_dct444   (quality > 90, 4:4:4) 8 threads per MCU: the 8x8 tile is loaded once, colour transform
//                   (stb:1552-1554), stb's float AAN DCT (stb:1240-1286) on rows, a shared-memory transpose, the same
//                   DCT on columns, quantisation + zigzag (stb:1314-1323) -> int16[3][64]; while the coefficients are
//                   still in shared memory the same threads compute the bit length of each block's AC symbols
//   k_jpeg_dcbits   adds the DC symbol (needs the previous block of the component) -> bit length per block
//   k_jpeg_dct + k_jpeg_entropy<false>   the same two steps for 4:2:0 (quality <= 90, chroma averaged as stb:1526-1535)
//   exclusive sum (rt_prims.cuh) -> bit offset of every block
//   k_jpeg_entropy<true>   one persistent warp per block, two zigzag positions per lane, zero runs from ballots of the
//                   non-zero mask (stb:1325-1365); the block's bits are assembled in shared memory and leave as words
//   k_jpeg_ffcount / scan / k_jpeg_stuff   the 0xFF -> 0xFF 0x00 byte stuffing of stb:1229-1232 as a
//                   count / scan / scatter of 32-byte chunks, the 1-padding of the last byte (stb:1567) and the EOI marker
// All float arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction) in stb's operation order, which
// is what a host compiler emits for the reference; the quantiser truncates like its (int) cast.
// Bound (ncu, profiles/r01_jpeg_ncu.md): instruction issue — k_jpeg_dct444 issues at 80 % of peak, k_jpeg_entropy<true> at
// 67 % with 23 of 32 lanes active — not HBM (3 B/pixel read, 6 B/pixel of int16 coefficients written and read once,
// ~1-2 B/pixel of stream: 6-11 % of the DRAM bandwidth).  Measured 1.13 ms for an 8K frame, 0.09 ms for 1200x600.
#include "rt_prims.cuh"

#include <cstring>
#include <vector>

#include "rt_jpeg.cuh"

namespace rtd {

namespace {

// ---- ITU-T T.81 Annex K: zigzag order, quantisation bases, typical Huffman specifications ----
const uint8_t kZigZag[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30,
                             41, 43, 9,  11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38,
                             46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};
const int kLumaQ[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,  14, 13, 16, 24, 40, 57,
                        69, 56, 14, 17, 22,  29,  51,  87,  80, 62, 18, 22, 37,  56,  68,  109, 103, 77, 24, 35, 55, 64,
                        81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const int kChromaQ[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                          99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                          99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
// number of codes of length 1..16, then the symbols in code order (tables K.3 - K.6)
const uint8_t kDcLumaBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChromaBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumaBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kAcChromaBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

// Device-side tables of one quality setting.  Huffman entries are code | length << 16 (stb keeps them as the
// precomputed arrays YDC_HT / YAC_HT / UVDC_HT / UVAC_HT, stb_image_write.h:1396-1433; here they are derived
// from the Annex K specifications above by the canonical construction, which yields the same codes).
struct JpegTables {
    float fdtbl[2][64]; // 1 / (quant * AAN scale), natural order (stb:1459-1464)
    uint32_t dc[2][12];
    uint32_t ac[2][256];
    uint8_t zigzag[64];
};

void canonical_codes(const uint8_t bits[16], const uint8_t* vals, uint32_t* table) {
    uint32_t code = 0;
    int k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < bits[len - 1]; ++i, ++k) table[vals[k]] = code++ | (uint32_t(len) << 16);
        code <<= 1;
    }
}

// quality mapping, quantisation tables and the JFIF/DQT/SOF0/DHT/SOS header of stb_image_write.h:1443-1491
void build_tables(int quality, int w, int h, JpegTables& t, std::vector<uint8_t>& header, bool& subsample) {
    memset(&t, 0, sizeof t);
    quality = quality ? quality : 90;
    subsample = quality <= 90;
    quality = quality < 1 ? 1 : quality > 100 ? 100 : quality;
    quality = quality < 50 ? 5000 / quality : 200 - quality * 2;
    uint8_t ytab[64], uvtab[64];
    for (int i = 0; i < 64; ++i) {
        int y = (kLumaQ[i] * quality + 50) / 100, c = (kChromaQ[i] * quality + 50) / 100;
        ytab[kZigZag[i]] = uint8_t(y < 1 ? 1 : y > 255 ? 255 : y);
        uvtab[kZigZag[i]] = uint8_t(c < 1 ? 1 : c > 255 ? 255 : c);
    }
    static const float aasf[8] = {1.0f * 2.828427125f,         1.387039845f * 2.828427125f, 1.306562965f * 2.828427125f,
                                  1.175875602f * 2.828427125f, 1.0f * 2.828427125f,         0.785694958f * 2.828427125f,
                                  0.541196100f * 2.828427125f, 0.275899379f * 2.828427125f};
    for (int row = 0, k = 0; row < 8; ++row)
        for (int col = 0; col < 8; ++col, ++k) {
            t.fdtbl[0][k] = 1 / (ytab[kZigZag[k]] * aasf[row] * aasf[col]);
            t.fdtbl[1][k] = 1 / (uvtab[kZigZag[k]] * aasf[row] * aasf[col]);
        }
    canonical_codes(kDcLumaBits, kDcVals, t.dc[0]);
    canonical_codes(kDcChromaBits, kDcVals, t.dc[1]);
    canonical_codes(kAcLumaBits, kAcLumaVals, t.ac[0]);
    canonical_codes(kAcChromaBits, kAcChromaVals, t.ac[1]);
    memcpy(t.zigzag, kZigZag, 64);

    header.clear();
    auto put = [&](std::initializer_list<int> b) {
        for (int v : b) header.push_back(uint8_t(v));
    };
    auto put_n = [&](const uint8_t* p, size_t n) { header.insert(header.end(), p, p + n); };
    put({0xFF, 0xD8, 0xFF, 0xE0, 0, 0x10, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0}); // SOI, APP0
    put({0xFF, 0xDB, 0, 0x84, 0});                                                            // DQT, table 0
    put_n(ytab, 64);
    put({1});
    put_n(uvtab, 64);
    put({0xFF, 0xC0, 0, 0x11, 8, h >> 8, h & 255, w >> 8, w & 255, 3, 1, subsample ? 0x22 : 0x11, 0, 2, 0x11, 1, 3, 0x11, 1});
    put({0xFF, 0xC4, 0x01, 0xA2, 0}); // DHT: luma DC
    put_n(kDcLumaBits, 16);
    put_n(kDcVals, 12);
    put({0x10}); // luma AC
    put_n(kAcLumaBits, 16);
    put_n(kAcLumaVals, 162);
    put({1}); // chroma DC
    put_n(kDcChromaBits, 16);
    put_n(kDcVals, 12);
    put({0x11}); // chroma AC
    put_n(kAcChromaBits, 16);
    put_n(kAcChromaVals, 162);
    put({0xFF, 0xDA, 0, 0xC, 3, 1, 0, 2, 0x11, 3, 0x11, 0, 0x3F, 0}); // SOS
}

// ---- pass 1: colour transform, DCT, quantisation -------------------------------------------------------------
// the 1-D float DCT of stb_image_write.h:1240-1286 (AAN), operation for operation, without contraction
__device__ __forceinline__ void dct8(float (&d)[8]) {
    const float tmp0 = __fadd_rn(d[0], d[7]), tmp7 = __fsub_rn(d[0], d[7]);
    const float tmp1 = __fadd_rn(d[1], d[6]), tmp6 = __fsub_rn(d[1], d[6]);
    const float tmp2 = __fadd_rn(d[2], d[5]), tmp5 = __fsub_rn(d[2], d[5]);
    const float tmp3 = __fadd_rn(d[3], d[4]), tmp4 = __fsub_rn(d[3], d[4]);
    // even part
    float tmp10 = __fadd_rn(tmp0, tmp3), tmp13 = __fsub_rn(tmp0, tmp3);
    float tmp11 = __fadd_rn(tmp1, tmp2), tmp12 = __fsub_rn(tmp1, tmp2);
    d[0] = __fadd_rn(tmp10, tmp11);
    d[4] = __fsub_rn(tmp10, tmp11);
    const float z1 = __fmul_rn(__fadd_rn(tmp12, tmp13), 0.707106781f);
    d[2] = __fadd_rn(tmp13, z1);
    d[6] = __fsub_rn(tmp13, z1);
    // odd part
    tmp10 = __fadd_rn(tmp4, tmp5);
    tmp11 = __fadd_rn(tmp5, tmp6);
    tmp12 = __fadd_rn(tmp6, tmp7);
    const float z5 = __fmul_rn(__fsub_rn(tmp10, tmp12), 0.382683433f);
    const float z2 = __fadd_rn(__fmul_rn(tmp10, 0.541196100f), z5);
    const float z4 = __fadd_rn(__fmul_rn(tmp12, 1.306562965f), z5);
    const float z3 = __fmul_rn(tmp11, 0.707106781f);
    const float z11 = __fadd_rn(tmp7, z3), z13 = __fsub_rn(tmp7, z3);
    d[5] = __fadd_rn(z13, z2);
    d[3] = __fsub_rn(z13, z2);
    d[1] = __fadd_rn(z11, z4);
    d[7] = __fsub_rn(z11, z4);
}

// component value of one pixel (stb_image_write.h:1513-1515), clamped to the image like stb's edge replication
__device__ __forceinline__ float pixel_comp(const uint8_t* __restrict__ rgb, int w, int h, int x, int y, int comp) {
    x = x < w ? x : w - 1;
    y = y < h ? y : h - 1;
    const uint8_t* p = rgb + (size_t(y) * w + x) * 3;
    const float r = p[0], g = p[1], b = p[2];
    if (comp == 0) return __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(0.29900f, r), __fmul_rn(0.58700f, g)), __fmul_rn(0.11400f, b)), 128.f);
    if (comp == 1) return __fadd_rn(__fsub_rn(__fmul_rn(-0.16874f, r), __fmul_rn(0.33126f, g)), __fmul_rn(0.50000f, b));
    return __fsub_rn(__fsub_rn(__fmul_rn(0.50000f, r), __fmul_rn(0.41869f, g)), __fmul_rn(0.08131f, b));
}

// block b -> component, pixel origin, subsampled?   4:4:4: MCU = (Y, U, V) of one 8x8 tile;
// 4:2:0 (stb:1498-1539): MCU = 4 Y blocks of a 16x16 tile + its averaged U and V
struct BlockGeom {
    int comp, x0, y0;
    bool sub;
};
__device__ __forceinline__ BlockGeom block_geom(uint32_t b, int mcus_x, bool subsample) {
    BlockGeom g;
    if (!subsample) {
        const uint32_t mcu = b / 3u;
        g.comp = int(b - mcu * 3u);
        g.x0 = int(mcu % uint32_t(mcus_x)) * 8;
        g.y0 = int(mcu / uint32_t(mcus_x)) * 8;
        g.sub = false;
    } else {
        const uint32_t mcu = b / 6u, k = b - mcu * 6u;
        g.x0 = int(mcu % uint32_t(mcus_x)) * 16;
        g.y0 = int(mcu / uint32_t(mcus_x)) * 16;
        g.comp = k < 4u ? 0 : int(k) - 3;
        g.sub = k >= 4u;
        if (k < 4u) {
            g.x0 += int(k & 1u) * 8;
            g.y0 += int(k >> 1) * 8;
        }
    }
    return g;
}
// index of the block whose DC predicts block b's (same component, previous in scan order), -1 for the first
__device__ __forceinline__ long long dc_pred_block(uint32_t b, bool subsample) {
    if (!subsample) return (long long)b - 3;
    const uint32_t k = b % 6u;
    if (k >= 1u && k <= 3u) return (long long)b - 1;
    return k == 0u ? (long long)b - 3 : (long long)b - 6;
}

#define JPG_DCT_THREADS 256
__global__ void __launch_bounds__(JPG_DCT_THREADS)
    k_jpeg_dct(const uint8_t* __restrict__ rgb, int w, int h, int mcus_x, int subsample, uint32_t n_blocks,
               const JpegTables* __restrict__ tab, int16_t* __restrict__ coef) {
    __shared__ float s_t[JPG_DCT_THREADS / 8][8][9];
    __shared__ __align__(16) int16_t s_q[JPG_DCT_THREADS / 8][64];
    __shared__ float s_fd[2][64];
    __shared__ uint8_t s_zz[64];
    if (threadIdx.x < 128) (&s_fd[0][0])[threadIdx.x] = (&tab->fdtbl[0][0])[threadIdx.x];
    if (threadIdx.x < 64) s_zz[threadIdx.x] = tab->zigzag[threadIdx.x];
    __syncthreads();
    const int g = threadIdx.x >> 3, r = threadIdx.x & 7;
    const uint32_t b = blockIdx.x * (JPG_DCT_THREADS / 8) + g;
    const bool live = b < n_blocks;
    float d[8];
    BlockGeom bg{0, 0, 0, false};
    if (live) {
        bg = block_geom(b, mcus_x, subsample != 0);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (!bg.sub) {
                d[c] = pixel_comp(rgb, w, h, bg.x0 + c, bg.y0 + r, bg.comp);
            } else { // (U[j] + U[j+1] + U[j+16] + U[j+17]) * 0.25f, stb:1531-1532
                const int x = bg.x0 + 2 * c, y = bg.y0 + 2 * r;
                float s = __fadd_rn(pixel_comp(rgb, w, h, x, y, bg.comp), pixel_comp(rgb, w, h, x + 1, y, bg.comp));
                s = __fadd_rn(s, pixel_comp(rgb, w, h, x, y + 1, bg.comp));
                s = __fadd_rn(s, pixel_comp(rgb, w, h, x + 1, y + 1, bg.comp));
                d[c] = __fmul_rn(s, 0.25f);
            }
        }
        dct8(d); // row r
#pragma unroll
        for (int c = 0; c < 8; ++c) s_t[g][r][c] = d[c];
    }
    __syncwarp();
    if (live) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = s_t[g][i][r];
        dct8(d); // column r: d[i] = coefficient (row i, column r)
        const int tq = bg.comp ? 1 : 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = i * 8 + r;
            const float v = __fmul_rn(d[i], s_fd[tq][j]);
            s_q[g][s_zz[j]] = int16_t(__float2int_rz(v < 0.f ? __fsub_rn(v, 0.5f) : __fadd_rn(v, 0.5f)));
        }
    }
    __syncwarp();
    if (live) reinterpret_cast<int4*>(coef + size_t(b) * 64)[r] = reinterpret_cast<const int4*>(s_q[g])[r];
}

// 4:4:4 (quality > 90, the reference's setting): the three blocks of an MCU share their pixels, so one group of 8
// threads loads the 8x8 tile ONCE (row r as six 32-bit words when the row is word-aligned) and runs the three
// transforms from registers; the group's 3 x 64 coefficients leave as 384 contiguous bytes.
__global__ void __launch_bounds__(JPG_DCT_THREADS)
    k_jpeg_dct444(const uint8_t* __restrict__ rgb, int w, int h, int mcus_x, uint32_t n_mcus, const JpegTables* __restrict__ tab,
                  int16_t* __restrict__ coef, uint32_t* __restrict__ ac_bits) {
    __shared__ float s_t[3][JPG_DCT_THREADS / 8][8][9];
    __shared__ __align__(16) int16_t s_q[JPG_DCT_THREADS / 8][3][64];
    __shared__ float s_fd[2][64];
    __shared__ uint8_t s_zz[64];
    __shared__ uint8_t s_aclen[2][256]; // code lengths of the AC symbols (the codes themselves are only needed when writing)
    if (threadIdx.x < 128) (&s_fd[0][0])[threadIdx.x] = (&tab->fdtbl[0][0])[threadIdx.x];
    if (threadIdx.x < 64) s_zz[threadIdx.x] = tab->zigzag[threadIdx.x];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&s_aclen[0][0])[i] = uint8_t((&tab->ac[0][0])[i] >> 16);
    __syncthreads();
    const int g = threadIdx.x >> 3, r = threadIdx.x & 7;
    const uint32_t mcu = blockIdx.x * (JPG_DCT_THREADS / 8) + g;
    const bool live = mcu < n_mcus;
    float d[3][8];
    if (live) {
        const int x0 = int(mcu % uint32_t(mcus_x)) * 8, y0 = int(mcu / uint32_t(mcus_x)) * 8;
        const int y = y0 + r < h ? y0 + r : h - 1; // edge rows / columns are replicated (stb:1546-1551)
        const uint8_t* row = rgb + (size_t(y) * w + x0) * 3;
        uint8_t px[24];
        if (x0 + 8 <= w && (reinterpret_cast<uintptr_t>(row) & 3u) == 0u) {
            const uint32_t* r4 = reinterpret_cast<const uint32_t*>(row);
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const uint32_t v = __ldg(r4 + k);
                px[4 * k] = uint8_t(v);
                px[4 * k + 1] = uint8_t(v >> 8);
                px[4 * k + 2] = uint8_t(v >> 16);
                px[4 * k + 3] = uint8_t(v >> 24);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int x = x0 + c < w ? c : w - 1 - x0;
#pragma unroll
                for (int k = 0; k < 3; ++k) px[3 * c + k] = row[3 * x + k];
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float R = px[3 * c], G = px[3 * c + 1], B = px[3 * c + 2];
            d[0][c] = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(0.29900f, R), __fmul_rn(0.58700f, G)), __fmul_rn(0.11400f, B)), 128.f);
            d[1][c] = __fadd_rn(__fsub_rn(__fmul_rn(-0.16874f, R), __fmul_rn(0.33126f, G)), __fmul_rn(0.50000f, B));
            d[2][c] = __fsub_rn(__fsub_rn(__fmul_rn(0.50000f, R), __fmul_rn(0.41869f, G)), __fmul_rn(0.08131f, B));
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            dct8(d[k]); // row r of component k
#pragma unroll
            for (int c = 0; c < 8; ++c) s_t[k][g][r][c] = d[k][c];
        }
    }
    __syncwarp();
    if (live) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int i = 0; i < 8; ++i) d[k][i] = s_t[k][g][i][r];
            dct8(d[k]); // column r: d[k][i] = coefficient (row i, column r)
            const int tq = k ? 1 : 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = i * 8 + r;
                const float v = __fmul_rn(d[k][i], s_fd[tq][j]);
                s_q[g][k][s_zz[j]] = int16_t(__float2int_rz(v < 0.f ? __fsub_rn(v, 0.5f) : __fadd_rn(v, 0.5f)));
            }
        }
    }
    __syncwarp();
    if (live) {
        int4* out = reinterpret_cast<int4*>(coef + size_t(mcu) * 192);
        const int4* in = reinterpret_cast<const int4*>(&s_q[g][0][0]);
#pragma unroll
        for (int k = 0; k < 3; ++k) out[k * 8 + r] = in[k * 8 + r];
    }
    // Bit length of the AC part of each block while its coefficients are still in shared memory (saves the separate
    // length pass over 6 B/pixel): thread r owns zigzag positions 8r .. 8r+7; the block's non-zero mask is assembled
    // with three butterfly steps inside the group of 8; runs and categories as in k_jpeg_entropy (stb:1336-1364).
    // All 32 lanes take part in the shuffles; dead groups (past the last MCU) carry zeros.
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int tq = k ? 1 : 0;
        int v[8];
        uint32_t m8 = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[j] = live ? int(s_q[g][k][8 * r + j]) : 0;
            if (v[j] != 0 && (r | j) != 0) m8 |= 1u << j; // position 0 is the DC coefficient
        }
        uint32_t lo = r < 4 ? m8 << (8 * r) : 0u, hi = r >= 4 ? m8 << (8 * (r - 4)) : 0u;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            lo |= __shfl_xor_sync(0xffffffffu, lo, o);
            hi |= __shfl_xor_sync(0xffffffffu, hi, o);
        }
        const unsigned long long mask = ((unsigned long long)hi << 32) | lo;
        const uint32_t zrl_len = s_aclen[tq][0xF0];
        uint32_t bits = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (m8 & (1u << j)) {
                const uint32_t pos = 8u * r + j;
                const unsigned long long below = mask & ((1ull << pos) - 1ull);
                const uint32_t prev = below ? 63u - uint32_t(__clzll((long long)below)) : 0u;
                const uint32_t run = pos - prev - 1u;
                const int a = v[j] < 0 ? -v[j] : v[j];
                const uint32_t nb = 32u - uint32_t(__clz(a));
                bits += (run >> 4) * zrl_len + s_aclen[tq][((run & 15u) << 4) + nb] + nb;
            }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
        if (live && r == 0) ac_bits[size_t(mcu) * 3 + k] = bits + ((mask >> 63) ? 0u : uint32_t(s_aclen[tq][0x00]));
    }
}

// DC part of the block lengths (stb:1325-1334): needs the previous block of the same component, so it runs after the DCT
__global__ void __launch_bounds__(256) k_jpeg_dcbits(const int16_t* __restrict__ coef, const uint32_t* __restrict__ ac_bits, uint32_t n_blocks,
                                                     const JpegTables* __restrict__ tab, unsigned long long* __restrict__ block_bits) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const int tq = (b % 3u) ? 1 : 0;
    const int diff = int(coef[size_t(b) * 64]) - (b >= 3u ? int(coef[size_t(b - 3u) * 64]) : 0);
    uint32_t n = 0u;
    if (diff != 0) {
        const int a = diff < 0 ? -diff : diff;
        n = 32u - uint32_t(__clz(a));
    }
    block_bits[b] = (unsigned long long)(ac_bits[b] + (tab->dc[tq][n] >> 16) + n);
}

// ---- pass 2/4: entropy coding ----------------------------------------------------------------------------------
// big-endian bit stream in 32-bit words: bit position p lives in word p >> 5 at bit 31 - (p & 31)
__device__ __forceinline__ void put_bits(uint32_t* __restrict__ words, unsigned long long pos, uint32_t code, uint32_t len) {
    if (len == 0u) return;
    const unsigned long long wi = pos >> 5;
    const uint32_t off = uint32_t(pos) & 31u;
    const unsigned long long v = (unsigned long long)code << (64u - off - len);
    atomicOr(words + wi, uint32_t(v >> 32));
    if (off + len > 32u) atomicOr(words + wi + 1, uint32_t(v));
}
// category and value bits of a coefficient (stb_image_write.h:1288-1296)
__device__ __forceinline__ void calc_bits(int val, uint32_t& bits, uint32_t& nbits) {
    const int a = val < 0 ? -val : val;
    nbits = 32u - uint32_t(__clz(a));
    bits = uint32_t(val < 0 ? val - 1 : val) & ((1u << nbits) - 1u);
}

#define JPG_ENT_THREADS 256
template <bool WRITE>
__global__ void __launch_bounds__(JPG_ENT_THREADS)
    k_jpeg_entropy(const int16_t* __restrict__ coef, uint32_t n_blocks, int subsample, const JpegTables* __restrict__ tab,
                   unsigned long long* __restrict__ block_bits, const unsigned long long* __restrict__ block_off,
                   uint32_t* __restrict__ words) {
    __shared__ uint32_t s_dc[2][12];
    __shared__ uint32_t s_ac[2][256];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&s_ac[0][0])[i] = (&tab->ac[0][0])[i];
    if (threadIdx.x < 24) (&s_dc[0][0])[threadIdx.x] = (&tab->dc[0][0])[threadIdx.x];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    // persistent warps: the 2 KB of code tables are staged once per CTA, not once per 8 blocks
    const uint32_t warps_total = gridDim.x * (JPG_ENT_THREADS / 32);
    for (uint32_t b = blockIdx.x * (JPG_ENT_THREADS / 32) + (threadIdx.x >> 5); b < n_blocks; b += warps_total) {
    const int16_t* c = coef + size_t(b) * 64;
    const int c_lo = c[lane], c_hi = c[lane + 32];
    const int tq = subsample ? ((b % 6u) >= 4u ? 1 : 0) : ((b % 3u) ? 1 : 0);

    // zero runs from the mask of non-zero AC positions (stb:1336-1361 walks them one by one)
    const uint32_t m_lo = __ballot_sync(0xffffffffu, lane != 0u && c_lo != 0);
    const uint32_t m_hi = __ballot_sync(0xffffffffu, c_hi != 0);
    const unsigned long long mask = ((unsigned long long)m_hi << 32) | m_lo;
    const uint32_t zrl = s_ac[tq][0xF0], eob = s_ac[tq][0x00];

    uint32_t len[2] = {0u, 0u}, code[2] = {0u, 0u}, nzrl[2] = {0u, 0u};
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t k = lane + 32u * half;
        const int v = half ? c_hi : c_lo;
        if (k == 0u) { // DC difference (stb:1325-1334)
            const long long pb = dc_pred_block(b, subsample != 0);
            const int diff = v - (pb >= 0 ? int(coef[size_t(pb) * 64]) : 0);
            if (diff == 0) {
                code[0] = s_dc[tq][0] & 0xffffu;
                len[0] = s_dc[tq][0] >> 16;
            } else {
                uint32_t bits, nb;
                calc_bits(diff, bits, nb);
                const uint32_t h = s_dc[tq][nb];
                code[0] = ((h & 0xffffu) << nb) | bits;
                len[0] = (h >> 16) + nb;
            }
        } else if (v != 0) {
            const unsigned long long below = mask & ((1ull << k) - 1ull);
            const uint32_t prev = below ? 63u - uint32_t(__clzll((long long)below)) : 0u;
            const uint32_t run = k - prev - 1u;
            uint32_t bits, nb;
            calc_bits(v, bits, nb);
            const uint32_t h = s_ac[tq][((run & 15u) << 4) + nb];
            nzrl[half] = run >> 4;
            code[half] = ((h & 0xffffu) << nb) | bits;
            len[half] = (h >> 16) + nb;
        }
    }
    const uint32_t tot_lo = len[0] + nzrl[0] * (zrl >> 16), tot_hi = len[1] + nzrl[1] * (zrl >> 16);
    // warp scans: offsets of positions 0..31, then 32..63
    uint32_t inc_lo = tot_lo, inc_hi = tot_hi;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, inc_lo, o), h = __shfl_up_sync(0xffffffffu, inc_hi, o);
        if (lane >= uint32_t(o)) {
            inc_lo += a;
            inc_hi += h;
        }
    }
    const uint32_t sum_lo = __shfl_sync(0xffffffffu, inc_lo, 31), sum_hi = __shfl_sync(0xffffffffu, inc_hi, 31);
    const bool need_eob = (mask >> 63) == 0ull; // last non-zero before position 63, or no AC at all (stb:1340-1343,1362-1364)
    const uint32_t total = sum_lo + sum_hi + (need_eob ? (eob >> 16) : 0u);
    if (!WRITE) {
        if (lane == 0u) block_bits[b] = total;
        continue;
    }
    // The block's bits are assembled in shared memory — aligned to the 32-bit words of the global stream — with shared
    // atomics; whole words then leave with plain stores and only the (at most two) words shared with the neighbouring
    // blocks need a global atomicOr.  (One global atomicOr per symbol cost 590 us on an 8K frame.)
    __shared__ uint32_t s_bits[JPG_ENT_THREADS / 32][58]; // 27 bits x 64 symbols + 31 bits of lead-in
    uint32_t* mine = s_bits[threadIdx.x >> 5];
    __syncwarp(); // the previous block's words have been read out
    mine[lane] = 0u;
    if (lane < 26u) mine[32u + lane] = 0u;
    __syncwarp();
    const unsigned long long base = block_off[b];
    const uint32_t lead = uint32_t(base) & 31u;
    auto put_local = [&](uint32_t pos, uint32_t cbits, uint32_t clen) { // pos relative to the block's first bit
        if (clen == 0u) return;
        const uint32_t q = lead + pos, wi = q >> 5, off = q & 31u;
        const unsigned long long v = (unsigned long long)cbits << (64u - off - clen);
        atomicOr(mine + wi, uint32_t(v >> 32));
        if (off + clen > 32u) atomicOr(mine + wi + 1, uint32_t(v));
    };
    uint32_t p = inc_lo - tot_lo;
    for (uint32_t z = 0; z < nzrl[0]; ++z, p += zrl >> 16) put_local(p, zrl & 0xffffu, zrl >> 16);
    put_local(p, code[0], len[0]);
    p = sum_lo + (inc_hi - tot_hi);
    for (uint32_t z = 0; z < nzrl[1]; ++z, p += zrl >> 16) put_local(p, zrl & 0xffffu, zrl >> 16);
    put_local(p, code[1], len[1]);
    if (need_eob && lane == 0u) put_local(sum_lo + sum_hi, eob & 0xffffu, eob >> 16);
    __syncwarp();
    const uint32_t end = lead + total, n_words = (end + 31u) >> 5;
    uint32_t* out = words + (base >> 5);
    for (uint32_t i = lane; i < n_words; i += 32u) {
        const bool shared_word = (i == 0u && lead != 0u) || (i == n_words - 1u && (end & 31u) != 0u);
        if (shared_word) atomicOr(out + i, mine[i]);
        else out[i] = mine[i];
    }
    } // persistent loop
}

// ---- pass 5: byte stuffing ---------------------------------------------------------------------------------------
#define JPG_STUFF_BYTES 32 // stream bytes per thread
// A thread's 32 stream bytes as 8 words in MEMORY order (byte k of the chunk = bits 8(k&3).. of S[k>>2]); the last
// byte of the stream is padded with ones (stb:1494,1567: fillBits = 7 ones, only whole bytes leave the bit buffer).
__device__ __forceinline__ void load_chunk(const uint32_t* __restrict__ words, uint32_t t, unsigned long long total_bits, uint32_t (&S)[8]) {
    const uint4* w4 = reinterpret_cast<const uint4*>(words) + size_t(t) * 2;
    const uint4 a = __ldg(w4), b = __ldg(w4 + 1);
    uint32_t W[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (total_bits & 7ull) {
        const unsigned long long last = total_bits >> 3; // index of the partial byte
        if ((last >> 5) == t) {
            const uint32_t k = uint32_t(last) & 31u;
            W[k >> 2] |= (0xFFu >> (total_bits & 7ull)) << (24u - 8u * (k & 3u));
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) S[k] = __byte_perm(W[k], 0u, 0x0123); // big-endian stream word -> memory order
}
__device__ __forceinline__ uint32_t count_ff(const uint32_t (&S)[8], uint32_t n_valid) {
    uint32_t n = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t eq = __vcmpeq4(S[k], 0xFFFFFFFFu); // 0xFF in every byte that is 0xFF
        if (4u * k + 4u > n_valid) eq &= n_valid > 4u * k ? (1u << (8u * (n_valid - 4u * k))) - 1u : 0u;
        n += uint32_t(__popc(eq)) >> 3;
    }
    return n;
}
__global__ void __launch_bounds__(256) k_jpeg_ffcount(const uint32_t* __restrict__ words, const unsigned long long* __restrict__ total_bits_p,
                                                      uint32_t n_threads, unsigned long long* __restrict__ ff) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    const unsigned long long total_bits = *total_bits_p, n_bytes = (total_bits + 7ull) >> 3;
    const unsigned long long i0 = (unsigned long long)t * JPG_STUFF_BYTES;
    const uint32_t n_valid = uint32_t(n_bytes - i0 < JPG_STUFF_BYTES ? (n_bytes > i0 ? n_bytes - i0 : 0ull) : JPG_STUFF_BYTES);
    uint32_t S[8];
    load_chunk(words, t, total_bits, S);
    ff[t] = count_ff(S, n_valid);
}
__global__ void __launch_bounds__(256) k_jpeg_stuff(const uint32_t* __restrict__ words, const unsigned long long* __restrict__ total_bits_p,
                                                    uint32_t n_threads, const unsigned long long* __restrict__ ff_off,
                                                    uint8_t* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    const unsigned long long total_bits = *total_bits_p, n_bytes = (total_bits + 7ull) >> 3;
    const unsigned long long i0 = (unsigned long long)t * JPG_STUFF_BYTES;
    const uint32_t n_valid = uint32_t(n_bytes - i0 < JPG_STUFF_BYTES ? (n_bytes > i0 ? n_bytes - i0 : 0ull) : JPG_STUFF_BYTES);
    uint32_t S[8];
    load_chunk(words, t, total_bits, S);
    uint8_t* o = out + i0 + ff_off[t];
    if (n_valid == JPG_STUFF_BYTES && ff_off[t + 1] == ff_off[t]) {
        // nothing to stuff (7 of 8 chunks): the 32 bytes move as aligned words, shifted to the destination's alignment
        const uint32_t head = (4u - (uint32_t(reinterpret_cast<uintptr_t>(o)) & 3u)) & 3u;
        if (head == 0u) {
            uint4* o4 = reinterpret_cast<uint4*>(o); // 4-byte aligned is all that is known: store word by word
            uint32_t* o1 = reinterpret_cast<uint32_t*>(o4);
#pragma unroll
            for (int k = 0; k < 8; ++k) o1[k] = S[k];
        } else {
            for (uint32_t k = 0; k < head; ++k) o[k] = uint8_t(S[0] >> (8u * k));
            uint32_t* o1 = reinterpret_cast<uint32_t*>(o + head);
#pragma unroll
            for (int k = 0; k < 7; ++k) o1[k] = __funnelshift_r(S[k], S[k + 1], 8u * head);
            for (uint32_t k = 0; k < 4u - head; ++k) o[head + 28u + k] = uint8_t(S[7] >> (8u * (head + k)));
        }
    } else {
        for (uint32_t k = 0; k < n_valid; ++k) {
            const uint32_t v = (S[k >> 2] >> (8u * (k & 3u))) & 255u;
            *o++ = uint8_t(v);
            if (v == 255u) *o++ = 0;
        }
        if (t == n_threads - 1u) { // the thread that holds the end of the stream: EOI (stb:1570-1571)
            o[0] = 0xFF;
            o[1] = 0xD9;
        }
    }
    if (t == n_threads - 1u && n_valid == JPG_STUFF_BYTES && ff_off[t + 1] == ff_off[t]) {
        o[32] = 0xFF;
        o[33] = 0xD9;
    }
}

template <class T>
bool grow(T*& p, size_t& cap, size_t need) {
    if (need <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t n = need + need / 4 + 64;
    if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return false;
    cap = n;
    return true;
}

} // namespace

struct JpegState {
    int sm_count = 0; // of the device the state was created on (launch sizing)
    JpegTables* d_tab = nullptr;
    int quality = -1;
    std::vector<uint8_t> header;
    bool subsample = false;
    int hdr_w = 0, hdr_h = 0;
    int16_t* coef = nullptr;
    size_t coef_cap = 0;
    uint32_t* ac_bits = nullptr; // [n_blocks] AC part of the lengths (4:4:4 path)
    size_t ac_bits_cap = 0;
    unsigned long long* bits = nullptr; // [n_blocks + 1] lengths, then (in place of a second array) ...
    size_t bits_cap = 0;
    unsigned long long* offs = nullptr; // [n_blocks + 1] exclusive sums; offs[n_blocks] = total bits
    size_t offs_cap = 0;
    uint32_t* words = nullptr;
    size_t words_cap = 0;
    unsigned long long* ff = nullptr; // [n_threads + 1] counts / exclusive sums
    size_t ff_cap = 0;
    unsigned long long* ff_off = nullptr;
    size_t ff_off_cap = 0;
    uint8_t* out = nullptr;
    size_t out_cap = 0;
    uint8_t* scan_tmp = nullptr;
    size_t scan_tmp_cap = 0;
    unsigned long long* h_pin = nullptr; // pinned [2]
    cudaEvent_t ev[2] = {nullptr, nullptr};
};

JpegState* jpeg_create() {
    JpegState* s = new JpegState();
    int dev = 0;
    bool ok = cudaGetDevice(&dev) == cudaSuccess &&
              cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d_tab, sizeof(JpegTables)) == cudaSuccess;
    ok = ok && cudaMallocHost(&s->h_pin, 2 * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaEventCreate(&s->ev[0]) == cudaSuccess && cudaEventCreate(&s->ev[1]) == cudaSuccess;
    if (!ok) {
        jpeg_destroy(s);
        return nullptr;
    }
    return s;
}

void jpeg_destroy(JpegState* s) {
    if (!s) return;
    cudaFree(s->d_tab);
    cudaFree(s->coef);
    cudaFree(s->bits);
    cudaFree(s->ac_bits);
    cudaFree(s->offs);
    cudaFree(s->words);
    cudaFree(s->ff);
    cudaFree(s->ff_off);
    cudaFree(s->out);
    cudaFree(s->scan_tmp);
    if (s->h_pin) cudaFreeHost(s->h_pin);
    for (auto& e : s->ev)
        if (e) cudaEventDestroy(e);
    delete s;
}

size_t jpeg_max_bytes(int w, int h) {
    // header + 3 (or 1.5) components x 64 coefficients x at most 27 bits, all of it stuffed, + EOI
    const size_t blocks = size_t((w + 7) / 8) * size_t((h + 7) / 8) * 3;
    return 1024 + blocks * 216 * 2 + 2;
}

#define JPG_TRY(x)                    \
    do {                              \
        cudaError_t e_ = (x);         \
        if (e_ != cudaSuccess) return e_; \
    } while (0)

cudaError_t jpeg_encode(JpegState* s, const uint8_t* rgb8_dev, int w, int h, int quality, uint8_t* out_host, size_t cap,
                        size_t* n_bytes, cudaStream_t st, float* ms_device) {
    *n_bytes = 0;
    if (quality != s->quality || w != s->hdr_w || h != s->hdr_h) {
        JpegTables t;
        build_tables(quality, w, h, t, s->header, s->subsample);
        JPG_TRY(cudaMemcpyAsync(s->d_tab, &t, sizeof t, cudaMemcpyHostToDevice, st));
        JPG_TRY(cudaStreamSynchronize(st)); // `t` lives on this stack frame
        s->quality = quality;
        s->hdr_w = w;
        s->hdr_h = h;
    }
    const int mcu = s->subsample ? 16 : 8;
    const int mcus_x = (w + mcu - 1) / mcu, mcus_y = (h + mcu - 1) / mcu;
    const size_t n_blocks = size_t(mcus_x) * mcus_y * (s->subsample ? 6 : 3);
    if (n_blocks >= (size_t(1) << 31)) return cudaErrorInvalidValue;
    if (!grow(s->coef, s->coef_cap, n_blocks * 64) || !grow(s->bits, s->bits_cap, n_blocks + 1) ||
        !grow(s->offs, s->offs_cap, n_blocks + 1))
        return cudaErrorMemoryAllocation;
    size_t tmp_bytes = prims::scan_scratch_elems(n_blocks + 1) * sizeof(unsigned long long);
    if (!grow(s->scan_tmp, s->scan_tmp_cap, tmp_bytes)) return cudaErrorMemoryAllocation;

    JPG_TRY(cudaEventRecord(s->ev[0], st));
    const unsigned dct_grid = unsigned((n_blocks + JPG_DCT_THREADS / 8 - 1) / (JPG_DCT_THREADS / 8));
    if (s->subsample) {
        k_jpeg_dct<<<dct_grid, JPG_DCT_THREADS, 0, st>>>(rgb8_dev, w, h, mcus_x, 1, uint32_t(n_blocks), s->d_tab, s->coef);
    } else {
        const uint32_t n_mcus = uint32_t(n_blocks / 3);
        if (!grow(s->ac_bits, s->ac_bits_cap, n_blocks)) return cudaErrorMemoryAllocation;
        k_jpeg_dct444<<<(n_mcus + JPG_DCT_THREADS / 8 - 1) / (JPG_DCT_THREADS / 8), JPG_DCT_THREADS, 0, st>>>(rgb8_dev, w, h, mcus_x, n_mcus,
                                                                                                   s->d_tab, s->coef, s->ac_bits);
    }
    unsigned ent_grid = unsigned((n_blocks + JPG_ENT_THREADS / 32 - 1) / (JPG_ENT_THREADS / 32));
    const unsigned ent_cap = unsigned(s->sm_count > 0 ? s->sm_count : 1) * 8u; // persistent warps, 8 CTAs of 256 threads per SM
    if (ent_grid > ent_cap) ent_grid = ent_cap;
    JPG_TRY(cudaMemsetAsync(s->bits + n_blocks, 0, sizeof(unsigned long long), st));
    if (s->subsample)
        k_jpeg_entropy<false><<<ent_grid, JPG_ENT_THREADS, 0, st>>>(s->coef, uint32_t(n_blocks), 1, s->d_tab, s->bits, nullptr, nullptr);
    else
        k_jpeg_dcbits<<<unsigned((n_blocks + 255) / 256), 256, 0, st>>>(s->coef, s->ac_bits, uint32_t(n_blocks), s->d_tab, s->bits);
    JPG_TRY(prims::exclusive_sum<unsigned long long>(s->bits, s->offs, n_blocks + 1, reinterpret_cast<unsigned long long*>(s->scan_tmp), st));
    // total bits -> host: sizes the word buffer and the stuffing grid
    JPG_TRY(cudaMemcpyAsync(s->h_pin, s->offs + n_blocks, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    JPG_TRY(cudaStreamSynchronize(st));
    const unsigned long long total_bits = s->h_pin[0];
    const size_t stream_bytes = size_t((total_bits + 7) >> 3);
    const size_t n_words = ((stream_bytes + 31) / 32) * 8 + 8; // whole 32-byte chunks for the stuffing pass + entropy slack
    if (!grow(s->words, s->words_cap, n_words)) return cudaErrorMemoryAllocation;
    JPG_TRY(cudaMemsetAsync(s->words, 0, n_words * sizeof(uint32_t), st));
    k_jpeg_entropy<true><<<ent_grid, JPG_ENT_THREADS, 0, st>>>(s->coef, uint32_t(n_blocks), s->subsample ? 1 : 0, s->d_tab, nullptr,
                                                               s->offs, s->words);
    const size_t n_threads = stream_bytes ? (stream_bytes + JPG_STUFF_BYTES - 1) / JPG_STUFF_BYTES : 1;
    if (!grow(s->ff, s->ff_cap, n_threads + 1) || !grow(s->ff_off, s->ff_off_cap, n_threads + 1)) return cudaErrorMemoryAllocation;
    const size_t tmp2 = prims::scan_scratch_elems(n_threads + 1) * sizeof(unsigned long long);
    if (!grow(s->scan_tmp, s->scan_tmp_cap, tmp2)) return cudaErrorMemoryAllocation;
    JPG_TRY(cudaMemsetAsync(s->ff + n_threads, 0, sizeof(unsigned long long), st));
    const unsigned st_grid = unsigned((n_threads + 255) / 256);
    k_jpeg_ffcount<<<st_grid, 256, 0, st>>>(s->words, s->offs + n_blocks, uint32_t(n_threads), s->ff);
    JPG_TRY(prims::exclusive_sum<unsigned long long>(s->ff, s->ff_off, n_threads + 1, reinterpret_cast<unsigned long long*>(s->scan_tmp), st));
    // worst case every byte is stuffed; the exact size comes back with the second sync
    const size_t hdr = s->header.size();
    if (!grow(s->out, s->out_cap, hdr + 2 * stream_bytes + 2)) return cudaErrorMemoryAllocation;
    JPG_TRY(cudaMemcpyAsync(s->out, s->header.data(), hdr, cudaMemcpyHostToDevice, st));
    k_jpeg_stuff<<<st_grid, 256, 0, st>>>(s->words, s->offs + n_blocks, uint32_t(n_threads), s->ff_off, s->out + hdr);
    JPG_TRY(cudaEventRecord(s->ev[1], st));
    JPG_TRY(cudaMemcpyAsync(s->h_pin + 1, s->ff_off + n_threads, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    JPG_TRY(cudaStreamSynchronize(st));
    JPG_TRY(cudaGetLastError());
    const size_t total = hdr + stream_bytes + size_t(s->h_pin[1]) + 2;
    *n_bytes = total;
    if (ms_device) JPG_TRY(cudaEventElapsedTime(ms_device, s->ev[0], s->ev[1]));
    if (!out_host) return cudaSuccess; // size query
    if (total > cap) return cudaErrorInvalidValue;
    JPG_TRY(cudaMemcpyAsync(out_host, s->out, total, cudaMemcpyDeviceToHost, st));
    JPG_TRY(cudaStreamSynchronize(st));
    return cudaSuccess;
}

} // namespace rtd
