// oracle/jpeg_decode_oracle.cpp — TEST INFRASTRUCTURE.  CPU restatement of the pixel stages of the reference's JPEG reader
// (vendored stb_image v2.26, src/libs/stb/stb_image.h; the reference calls stbi_loadf, src/main.cu:376-380): from
// decoded DCT coefficients to bytes — dequantisation (stb:2998-3003), integer IDCT (stb:2358-2447), row selection and
// up-sampling (stb:3355-3426, 3545-3555, 3800-3841), YCbCr -> RGB (stb:3558-3583).  Written row by row the way stb's
// load_jpeg_image walks the picture.  The coefficients come from the product's host half (rt_jpeg_parse), which has no
// oracle of its own: parse + these stages together are pinned byte-exact against the real stb (oracle/_ref/libref_stb.so,
// tests/test_jpeg_decode.py) and against committed stb outputs (tests/golden/jpeg_decode_golden.npz).
// Only tests/ may use this file.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/rt_api.h"

namespace {

inline int f2f(double x) { return int(x * 4096 + 0.5); }

void idct_1d(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7, int (&x)[4], int (&t)[4]) {
    int p2 = s2, p3 = s6;
    int p1 = (p2 + p3) * f2f(0.5411961f);
    int t2 = p1 + p3 * f2f(-1.847759065f);
    int t3 = p1 + p2 * f2f(0.765366865f);
    p2 = s0;
    p3 = s4;
    int t0 = (p2 + p3) * 4096, t1 = (p2 - p3) * 4096;
    x[0] = t0 + t3;
    x[3] = t0 - t3;
    x[1] = t1 + t2;
    x[2] = t1 - t2;
    t0 = s7;
    t1 = s5;
    t2 = s3;
    t3 = s1;
    p3 = t0 + t2;
    int p4 = t1 + t3;
    p1 = t0 + t3;
    p2 = t1 + t2;
    int p5 = (p3 + p4) * f2f(1.175875602f);
    t0 = t0 * f2f(0.298631336f);
    t1 = t1 * f2f(2.053119869f);
    t2 = t2 * f2f(3.072711026f);
    t3 = t3 * f2f(1.501321110f);
    p1 = p5 + p1 * f2f(-0.899976223f);
    p2 = p5 + p2 * f2f(-2.562915447f);
    p3 = p3 * f2f(-1.961570560f);
    p4 = p4 * f2f(-0.390180644f);
    t[3] = t3 + p1 + p4;
    t[2] = t2 + p2 + p3;
    t[1] = t1 + p2 + p4;
    t[0] = t0 + p1 + p3;
}
inline uint8_t clamp8(int v) { return uint8_t(v < 0 ? 0 : v > 255 ? 255 : v); }

void idct_block(uint8_t* out, int stride, const int16_t* coef, const uint16_t* dq) {
    int d[64], val[64];
    for (int i = 0; i < 64; ++i) d[i] = int16_t(coef[i] * dq[i]);
    for (int i = 0; i < 8; ++i) {
        int x[4], t[4];
        idct_1d(d[i], d[8 + i], d[16 + i], d[24 + i], d[32 + i], d[40 + i], d[48 + i], d[56 + i], x, t);
        for (int& v : x) v += 512;
        val[i] = (x[0] + t[3]) >> 10;
        val[56 + i] = (x[0] - t[3]) >> 10;
        val[8 + i] = (x[1] + t[2]) >> 10;
        val[48 + i] = (x[1] - t[2]) >> 10;
        val[16 + i] = (x[2] + t[1]) >> 10;
        val[40 + i] = (x[2] - t[1]) >> 10;
        val[24 + i] = (x[3] + t[0]) >> 10;
        val[32 + i] = (x[3] - t[0]) >> 10;
    }
    for (int i = 0; i < 8; ++i, out += stride) {
        const int* v = val + 8 * i;
        int x[4], t[4];
        idct_1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], x, t);
        for (int& q : x) q += 65536 + (128 << 17);
        out[0] = clamp8((x[0] + t[3]) >> 17);
        out[7] = clamp8((x[0] - t[3]) >> 17);
        out[1] = clamp8((x[1] + t[2]) >> 17);
        out[6] = clamp8((x[1] - t[2]) >> 17);
        out[2] = clamp8((x[2] + t[1]) >> 17);
        out[5] = clamp8((x[2] - t[1]) >> 17);
        out[3] = clamp8((x[3] + t[0]) >> 17);
        out[4] = clamp8((x[3] - t[0]) >> 17);
    }
}

// the four filters of stb_image.h:3355-3426 and the nearest-neighbour fallback :3545-3555, whole rows
void resample_row(uint8_t* out, const uint8_t* nr, const uint8_t* fr, int w, int hs, int vs) {
    if (hs == 1 && vs == 1) {
        memcpy(out, nr, size_t(w));
    } else if (hs == 1 && vs == 2) {
        for (int i = 0; i < w; ++i) out[i] = uint8_t((3 * nr[i] + fr[i] + 2) >> 2);
    } else if (hs == 2 && vs == 1) {
        if (w == 1) {
            out[0] = out[1] = nr[0];
            return;
        }
        out[0] = nr[0];
        out[1] = uint8_t((nr[0] * 3 + nr[1] + 2) >> 2);
        int i;
        for (i = 1; i < w - 1; ++i) {
            const int n = 3 * nr[i] + 2;
            out[i * 2] = uint8_t((n + nr[i - 1]) >> 2);
            out[i * 2 + 1] = uint8_t((n + nr[i + 1]) >> 2);
        }
        out[i * 2] = uint8_t((nr[w - 2] * 3 + nr[w - 1] + 2) >> 2);
        out[i * 2 + 1] = nr[w - 1];
    } else if (hs == 2 && vs == 2) {
        if (w == 1) {
            out[0] = out[1] = uint8_t((3 * nr[0] + fr[0] + 2) >> 2);
            return;
        }
        int t1 = 3 * nr[0] + fr[0];
        out[0] = uint8_t((t1 + 2) >> 2);
        for (int i = 1; i < w; ++i) {
            const int t0 = t1;
            t1 = 3 * nr[i] + fr[i];
            out[i * 2 - 1] = uint8_t((3 * t0 + t1 + 8) >> 4);
            out[i * 2] = uint8_t((3 * t1 + t0 + 8) >> 4);
        }
        out[w * 2 - 1] = uint8_t((t1 + 2) >> 2);
    } else {
        for (int i = 0; i < w; ++i)
            for (int j = 0; j < hs; ++j) out[i * hs + j] = nr[i];
    }
}

inline int fix(float x) { return int(x * 4096.0f + 0.5f) << 8; }

} // namespace

// coefficient planes (rt_jpeg_parse) -> bytes: height rows of width pixels with 3 channels (1 for a one-component file)
extern "C" int orc_jpeg_pixels(const rt_jpeg_coefficients* c, uint8_t* out) {
    if (!c || !out || c->n_comp < 1 || c->n_comp > 3) return 1;
    std::vector<std::vector<uint8_t>> plane(c->n_comp), line(c->n_comp);
    struct State {
        int hs, vs, ystep, w_lores, ypos, line0, line1;
    } st[3];
    for (int k = 0; k < c->n_comp; ++k) {
        const rt_jpeg_component& p = c->comp[k];
        plane[k].assign(size_t(p.w2) * p.h2, 0);
        const int bw = (p.x + 7) >> 3, bh = (p.y + 7) >> 3;
        for (int j = 0; j < bh; ++j)
            for (int i = 0; i < bw; ++i)
                idct_block(plane[k].data() + size_t(p.w2) * j * 8 + i * 8, p.w2, p.coeff + 64 * size_t(i + j * p.blocks_w), c->dequant[p.tq]);
        line[k].assign(size_t(c->width) + 8, 0);
        st[k].hs = c->h_max / p.h;
        st[k].vs = c->v_max / p.v;
        st[k].ystep = st[k].vs >> 1;
        st[k].w_lores = (c->width + st[k].hs - 1) / st[k].hs;
        st[k].ypos = 0;
        st[k].line0 = st[k].line1 = 0;
    }
    const int n = c->n_comp == 1 ? 1 : 3;
    for (int j = 0; j < c->height; ++j) {
        for (int k = 0; k < c->n_comp; ++k) { // stb_image.h:3828-3841
            State& r = st[k];
            const rt_jpeg_component& p = c->comp[k];
            const bool y_bot = r.ystep >= (r.vs >> 1);
            const uint8_t* l0 = plane[k].data() + size_t(r.line0) * p.w2;
            const uint8_t* l1 = plane[k].data() + size_t(r.line1) * p.w2;
            resample_row(line[k].data(), y_bot ? l1 : l0, y_bot ? l0 : l1, r.w_lores, r.hs, r.vs);
            if (++r.ystep >= r.vs) {
                r.ystep = 0;
                r.line0 = r.line1;
                if (++r.ypos < p.y) r.line1 += 1;
            }
        }
        uint8_t* o = out + size_t(j) * c->width * n;
        for (int i = 0; i < c->width; ++i, o += n) {
            if (n == 1) {
                o[0] = line[0][i];
            } else if (c->is_rgb) {
                o[0] = line[0][i];
                o[1] = line[1][i];
                o[2] = line[2][i];
            } else { // stbi__YCbCr_to_RGB_row (stb_image.h:3558-3583)
                const int y_fixed = (line[0][i] << 20) + (1 << 19);
                const int cr = line[2][i] - 128, cb = line[1][i] - 128;
                int r = y_fixed + cr * fix(1.40200f);
                int g = y_fixed + (cr * -fix(0.71414f)) + ((cb * -fix(0.34414f)) & 0xffff0000);
                int b = y_fixed + cb * fix(1.77200f);
                o[0] = clamp8(r >> 20);
                o[1] = clamp8(g >> 20);
                o[2] = clamp8(b >> 20);
            }
        }
    }
    return 0;
}
