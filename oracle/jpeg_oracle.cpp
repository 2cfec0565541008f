// oracle/jpeg_oracle.cpp — TEST INFRASTRUCTURE.  CPU restatement of the reference's JPEG writer: the vendored
// stb_image_write v1.15 (src/libs/stb/stb_image_write.h:1220-1573; Jon Olick's baseline encoder), which the
// reference calls as stbi_write_jpg("render.jpg", W, H, 3, data, 100) (src/main.cu:491).
// PINNED: byte-exact against the real stb compiled from /root/reference where it lies (oracle/_ref/libref_stb.so,
// tests/test_oracle_pin_live.py) and against the committed outputs of that build (tests/golden/jpeg_golden.npz,
// made by tests/golden/make_jpeg_golden.py).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file; the product
// (raytracing_renderer_cuda_b200/csrc/rt_jpeg.cu) never does.
//
// Written as the plain sequential algorithm: one MCU after the other, symbols appended to a byte vector through a
// bit accumulator.  Compile with -ffp-contract=off: stb's float DCT is evaluated without fused multiply-adds.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

// ITU-T T.81 Annex K (stb_image_write.h:1220-1221, 1370-1395, 1434-1437)
const int ZZ[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30,
                    41, 43, 9,  11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38,
                    46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};
const int QY[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,  14, 13, 16, 24, 40, 57,
                    69, 56, 14, 17, 22,  29,  51,  87,  80, 62, 18, 22, 37,  56,  68,  109, 103, 77, 24, 35, 55, 64,
                    81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const int QC[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                    99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
struct HuffSpec {
    uint8_t bits[16];
    std::vector<uint8_t> vals;
};
std::vector<uint8_t> seq12() {
    std::vector<uint8_t> v;
    for (int i = 0; i < 12; ++i) v.push_back(uint8_t(i));
    return v;
}
// AC symbol order of tables K.5 / K.6, written as (run << 4 | size) rows
const uint8_t AC_Y[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81,
    0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18,
    0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75,
    0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99,
    0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5,
    0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t AC_C[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08,
    0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25,
    0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47,
    0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74,
    0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97,
    0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4,
    0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const HuffSpec DC_Y_SPEC = {{0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0}, seq12()};
const HuffSpec DC_C_SPEC = {{0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0}, seq12()};
const HuffSpec AC_Y_SPEC = {{0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d}, std::vector<uint8_t>(AC_Y, AC_Y + 162)};
const HuffSpec AC_C_SPEC = {{0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77}, std::vector<uint8_t>(AC_C, AC_C + 162)};

struct Code {
    uint16_t code, len;
};
// stb stores these as literal tables (stb_image_write.h:1396-1433); they are the canonical codes of the specs
void make_codes(const HuffSpec& s, Code out[256]) {
    memset(out, 0, 256 * sizeof(Code));
    unsigned code = 0;
    size_t k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < s.bits[len - 1]; ++i) out[s.vals[k++]] = Code{uint16_t(code++), uint16_t(len)};
        code <<= 1;
    }
}

struct Writer { // stbiw__jpg_writeBits (stb_image_write.h:1223-1238)
    std::vector<uint8_t>& out;
    int buf = 0, cnt = 0;
    void put(Code c) {
        cnt += c.len;
        buf |= int(c.code) << (24 - cnt);
        while (cnt >= 8) {
            uint8_t b = uint8_t((buf >> 16) & 255);
            out.push_back(b);
            if (b == 255) out.push_back(0);
            buf <<= 8;
            cnt -= 8;
        }
    }
};

void dct1d(float* p, int stride) { // stbiw__jpg_DCT (stb_image_write.h:1240-1286)
    float d0 = p[0], d1 = p[stride], d2 = p[2 * stride], d3 = p[3 * stride], d4 = p[4 * stride], d5 = p[5 * stride],
          d6 = p[6 * stride], d7 = p[7 * stride];
    float t0 = d0 + d7, t7 = d0 - d7, t1 = d1 + d6, t6 = d1 - d6, t2 = d2 + d5, t5 = d2 - d5, t3 = d3 + d4, t4 = d3 - d4;
    float t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    p[0] = t10 + t11;
    p[4 * stride] = t10 - t11;
    float z1 = (t12 + t13) * 0.707106781f;
    p[2 * stride] = t13 + z1;
    p[6 * stride] = t13 - z1;
    t10 = t4 + t5;
    t11 = t5 + t6;
    t12 = t6 + t7;
    float z5 = (t10 - t12) * 0.382683433f;
    float z2 = t10 * 0.541196100f + z5;
    float z4 = t12 * 1.306562965f + z5;
    float z3 = t11 * 0.707106781f;
    float z11 = t7 + z3, z13 = t7 - z3;
    p[5 * stride] = z13 + z2;
    p[3 * stride] = z13 - z2;
    p[stride] = z11 + z4;
    p[7 * stride] = z11 - z4;
}

Code value_bits(int v) { // stbiw__jpg_calcBits (stb_image_write.h:1288-1296)
    int a = v < 0 ? -v : v;
    v = v < 0 ? v - 1 : v;
    uint16_t n = 1;
    while (a >>= 1) ++n;
    return Code{uint16_t(v & ((1 << n) - 1)), n};
}

// stbiw__jpg_processDU (stb_image_write.h:1298-1366): returns the quantised DC
int encode_block(Writer& w, float* du, int stride, const float* fdtbl, int dc_pred, const Code* dc, const Code* ac) {
    for (int r = 0; r < 8; ++r) dct1d(du + r * stride, 1);
    for (int c = 0; c < 8; ++c) dct1d(du + c, stride);
    int q[64];
    for (int y = 0, j = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x, ++j) {
            float v = du[y * stride + x] * fdtbl[j];
            q[ZZ[j]] = int(v < 0 ? v - 0.5f : v + 0.5f);
        }
    int diff = q[0] - dc_pred;
    if (diff == 0) {
        w.put(dc[0]);
    } else {
        Code b = value_bits(diff);
        w.put(dc[b.len]);
        w.put(b);
    }
    int last = 63;
    while (last > 0 && q[last] == 0) --last;
    if (last == 0) {
        w.put(ac[0x00]);
        return q[0];
    }
    for (int i = 1; i <= last; ++i) {
        int start = i;
        while (q[i] == 0 && i <= last) ++i;
        int run = i - start;
        for (int k = 0; k < (run >> 4); ++k) w.put(ac[0xF0]);
        run &= 15;
        Code b = value_bits(q[i]);
        w.put(ac[(run << 4) + b.len]);
        w.put(b);
    }
    if (last != 63) w.put(ac[0x00]);
    return q[0];
}

} // namespace

// stbi_write_jpg_core (stb_image_write.h:1368-1573) for comp == 3.  Returns the file size (0 if cap is too small).
extern "C" size_t orc_jpeg_encode(const uint8_t* rgb, int width, int height, int quality, uint8_t* out, size_t cap) {
    if (!rgb || width <= 0 || height <= 0) return 0;
    quality = quality ? quality : 90;
    const bool subsample = quality <= 90;
    quality = quality < 1 ? 1 : quality > 100 ? 100 : quality;
    quality = quality < 50 ? 5000 / quality : 200 - quality * 2;
    uint8_t ty[64], tc[64];
    for (int i = 0; i < 64; ++i) {
        int y = (QY[i] * quality + 50) / 100, c = (QC[i] * quality + 50) / 100;
        ty[ZZ[i]] = uint8_t(y < 1 ? 1 : y > 255 ? 255 : y);
        tc[ZZ[i]] = uint8_t(c < 1 ? 1 : c > 255 ? 255 : c);
    }
    static const float aasf[8] = {1.0f * 2.828427125f,         1.387039845f * 2.828427125f, 1.306562965f * 2.828427125f,
                                  1.175875602f * 2.828427125f, 1.0f * 2.828427125f,         0.785694958f * 2.828427125f,
                                  0.541196100f * 2.828427125f, 0.275899379f * 2.828427125f};
    float fy[64], fc[64];
    for (int row = 0, k = 0; row < 8; ++row)
        for (int col = 0; col < 8; ++col, ++k) {
            fy[k] = 1 / (ty[ZZ[k]] * aasf[row] * aasf[col]);
            fc[k] = 1 / (tc[ZZ[k]] * aasf[row] * aasf[col]);
        }
    Code dcy[256], dcc[256], acy[256], acc[256];
    make_codes(DC_Y_SPEC, dcy);
    make_codes(DC_C_SPEC, dcc);
    make_codes(AC_Y_SPEC, acy);
    make_codes(AC_C_SPEC, acc);

    std::vector<uint8_t> f;
    auto bytes = [&](std::initializer_list<int> l) {
        for (int v : l) f.push_back(uint8_t(v));
    };
    auto spec = [&](int info, const HuffSpec& s) {
        f.push_back(uint8_t(info));
        f.insert(f.end(), s.bits, s.bits + 16);
        f.insert(f.end(), s.vals.begin(), s.vals.end());
    };
    bytes({0xFF, 0xD8, 0xFF, 0xE0, 0, 0x10, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0, 0xFF, 0xDB, 0, 0x84, 0});
    f.insert(f.end(), ty, ty + 64);
    f.push_back(1);
    f.insert(f.end(), tc, tc + 64);
    bytes({0xFF, 0xC0, 0, 0x11, 8, height >> 8, height & 255, width >> 8, width & 255, 3, 1, subsample ? 0x22 : 0x11, 0, 2, 0x11, 1, 3,
           0x11, 1, 0xFF, 0xC4, 0x01, 0xA2});
    spec(0x00, DC_Y_SPEC);
    spec(0x10, AC_Y_SPEC);
    spec(0x01, DC_C_SPEC);
    spec(0x11, AC_C_SPEC);
    bytes({0xFF, 0xDA, 0, 0xC, 3, 1, 0, 2, 0x11, 3, 0x11, 0, 0x3F, 0});

    Writer w{f};
    int dc_y = 0, dc_u = 0, dc_v = 0;
    const int step = subsample ? 16 : 8;
    std::vector<float> Y(step * step), U(step * step), V(step * step);
    for (int y0 = 0; y0 < height; y0 += step)
        for (int x0 = 0; x0 < width; x0 += step) {
            for (int r = 0, pos = 0; r < step; ++r) {
                const int yy = y0 + r < height ? y0 + r : height - 1; // edge rows/columns are replicated
                for (int c = 0; c < step; ++c, ++pos) {
                    const int xx = x0 + c < width ? x0 + c : width - 1;
                    const uint8_t* p = rgb + (size_t(yy) * width + xx) * 3;
                    float R = p[0], G = p[1], B = p[2];
                    Y[pos] = +0.29900f * R + 0.58700f * G + 0.11400f * B - 128;
                    U[pos] = -0.16874f * R - 0.33126f * G + 0.50000f * B;
                    V[pos] = +0.50000f * R - 0.41869f * G - 0.08131f * B;
                }
            }
            if (subsample) {
                dc_y = encode_block(w, Y.data() + 0, 16, fy, dc_y, dcy, acy);
                dc_y = encode_block(w, Y.data() + 8, 16, fy, dc_y, dcy, acy);
                dc_y = encode_block(w, Y.data() + 128, 16, fy, dc_y, dcy, acy);
                dc_y = encode_block(w, Y.data() + 136, 16, fy, dc_y, dcy, acy);
                float su[64], sv[64];
                for (int yy = 0, pos = 0; yy < 8; ++yy)
                    for (int xx = 0; xx < 8; ++xx, ++pos) {
                        const int j = yy * 32 + xx * 2;
                        su[pos] = (U[j] + U[j + 1] + U[j + 16] + U[j + 17]) * 0.25f;
                        sv[pos] = (V[j] + V[j + 1] + V[j + 16] + V[j + 17]) * 0.25f;
                    }
                dc_u = encode_block(w, su, 8, fc, dc_u, dcc, acc);
                dc_v = encode_block(w, sv, 8, fc, dc_v, dcc, acc);
            } else {
                dc_y = encode_block(w, Y.data(), 8, fy, dc_y, dcy, acy);
                dc_u = encode_block(w, U.data(), 8, fc, dc_u, dcc, acc);
                dc_v = encode_block(w, V.data(), 8, fc, dc_v, dcc, acc);
            }
        }
    w.put(Code{0x7F, 7}); // pad the last byte with ones
    f.push_back(0xFF);
    f.push_back(0xD9);
    if (f.size() > cap || !out) return out ? 0 : f.size();
    memcpy(out, f.data(), f.size());
    return f.size();
}
