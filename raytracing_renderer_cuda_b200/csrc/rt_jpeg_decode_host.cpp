// rt_jpeg_decode_host.cpp — marker parsing and Huffman decoding of a JPEG file on the host (see the header).
// Follows the decoder the reference uses, stb_image v2.26 (vendored at src/libs/stb/stb_image.h; cited as "stb:"), so that
// the coefficients — and with the device stages of rt_jpeg_decode.cu the pixels — are the ones stbi_loadf produces
// (main.cu:376-380).  Codes of up to 9 bits are found through a look-up table like stb's (stb:1986-2000), longer ones by
// the canonical maxcode search (stb:2060-2084); stb's combined run/value table for small AC coefficients (stb:2002-2028) is
// an acceleration with identical results and is not reproduced.
#include "rt_jpeg_decode_host.hpp"

#include <cstring>
#include <exception>

namespace rtj {

namespace {

const uint8_t kDezigzag[64 + 15] = { // position in the zigzag stream -> row-major index (stb:2121-2136)
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
    6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
    39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct Huffman { // stb:1875-1886, 1942-1984
    uint8_t values[256] = {};
    uint8_t size[257] = {};
    uint32_t maxcode[18] = {};
    int delta[17] = {};
    uint16_t code[256] = {};
    uint8_t fast[512] = {}; // top 9 bits of the stream -> index into values[], 255 = longer code
    bool defined = false;   // a DHT segment has filled this table (a scan that selects an undefined one is rejected)
    Huffman() {
        memset(fast, 255, sizeof fast);
        maxcode[17] = 0xffffffffu; // the sentinel of the code-length search, whatever happens to the table later
    }
    bool build(const int* count) {
        defined = false;
        int k = 0;
        for (int i = 0; i < 16; ++i)
            for (int j = 0; j < count[i]; ++j) {
                if (k >= 256) return false;
                size[k++] = uint8_t(i + 1);
            }
        size[k] = 0;
        unsigned code = 0;
        k = 0;
        int j;
        for (j = 1; j <= 16; ++j) {
            delta[j] = k - int(code);
            if (size[k] == j) {
                while (size[k] == j) this->code[k++] = uint16_t(code++);
                if (code - 1 >= (1u << j)) return false;
            }
            maxcode[j] = code << (16 - j);
            code <<= 1;
        }
        maxcode[j] = 0xffffffffu;
        memset(fast, 255, sizeof fast);
        for (int i = 0; i < k; ++i) {
            const int sz = size[i];
            if (sz <= 9) {
                const int c = this->code[i] << (9 - sz), m = 1 << (9 - sz);
                for (int q = 0; q < m; ++q) fast[c + q] = uint8_t(i);
            }
        }
        defined = true;
        return true;
    }
};

const uint32_t kMask[17] = {0, 1, 3, 7, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383, 32767, 65535};
const int kBias[16] = {0, -1, -3, -7, -15, -31, -63, -127, -255, -511, -1023, -2047, -4095, -8191, -16383, -32767};
const uint8_t kNoMarker = 0xff;

struct Decoder {
    const uint8_t* p;
    const uint8_t* end;
    CoefficientImage& img;
    Huffman huff_dc[4], huff_ac[4];
    // entropy-decoder state (stb:1922-1936)
    uint32_t code_buffer = 0;
    int code_bits = 0;
    uint8_t marker = kNoMarker;
    bool nomore = false;
    int spec_start = 0, spec_end = 0, succ_high = 0, succ_low = 0, eob_run = 0;
    int scan_n = 0, order[4] = {0, 0, 0, 0};
    int restart_interval = 0, todo = 0;
    int mcu_x = 0, mcu_y = 0;
    bool jfif = false;
    int app14 = -1, rgb_ids = 0;

    Decoder(const uint8_t* d, size_t n, CoefficientImage& o) : p(d), end(d + n), img(o) {}

    bool fail(const char* why) {
        img.error = why;
        return false;
    }
    int get8() { return p < end ? *p++ : 0; } // stb returns 0 past the end of the data
    int get16() {
        const int hi = get8();
        return (hi << 8) | get8();
    }
    void skip(int n) { p = (end - p) < n ? end : p + n; }
    bool at_eof() const { return p >= end; }

    uint8_t get_marker() { // stb:2845-2854
        if (marker != kNoMarker) {
            const uint8_t x = marker;
            marker = kNoMarker;
            return x;
        }
        int x = get8();
        if (x != 0xff) return kNoMarker;
        while (x == 0xff) x = get8();
        return uint8_t(x);
    }

    void grow() { // stb:2011-2027: byte-wise refill; 0xFF00 -> 0xFF, any other marker stops the scan (zeros follow)
        do {
            unsigned b = nomore ? 0u : unsigned(get8());
            if (b == 0xff) {
                int c = get8();
                while (c == 0xff) c = get8();
                if (c != 0) {
                    marker = uint8_t(c);
                    nomore = true;
                    return;
                }
            }
            code_buffer |= b << (24 - code_bits);
            code_bits += 8;
        } while (code_bits <= 24);
    }
    int huff_decode(const Huffman& h) { // stb:2033-2085
        if (code_bits < 16) grow();
        const int f = h.fast[code_buffer >> 23];
        if (f < 255) {
            const int sz = h.size[f];
            if (sz > code_bits) return -1;
            code_buffer <<= sz;
            code_bits -= sz;
            return h.values[f];
        }
        const uint32_t temp = code_buffer >> 16;
        int k;
        for (k = 10;; ++k)
            if (temp < h.maxcode[k]) break;
        if (k == 17) {
            code_bits -= 16;
            return -1;
        }
        if (k > code_bits) return -1;
        const int c = int((code_buffer >> (32 - k)) & kMask[k]) + h.delta[k];
        if (c < 0 || c > 255) return -1;
        code_bits -= k;
        code_buffer <<= k;
        return h.values[c];
    }
    static uint32_t rotl(uint32_t x, int n) { return n ? (x << n) | (x >> (32 - n)) : x; }
    int extend_receive(int n) { // stb:2089-2103
        if (code_bits < n) grow();
        const int sgn = int32_t(code_buffer) >> 31;
        uint32_t k = rotl(code_buffer, n);
        if (n < 0 || n > 16) return 0;
        code_buffer = k & ~kMask[n];
        k &= kMask[n];
        code_bits -= n;
        return int(k) + (n < 16 ? (kBias[n] & ~sgn) : 0);
    }
    int get_bits(int n) { // stb:2106-2115
        if (code_bits < n) grow();
        uint32_t k = rotl(code_buffer, n);
        code_buffer = k & ~kMask[n];
        k &= kMask[n];
        code_bits -= n;
        return int(k);
    }
    bool get_bit() { // stb:2117-2125
        if (code_bits < 1) grow();
        const uint32_t k = code_buffer;
        code_buffer <<= 1;
        --code_bits;
        return (k & 0x80000000u) != 0;
    }

    void reset() { // stb:2862-2873
        code_bits = 0;
        code_buffer = 0;
        nomore = false;
        for (auto& c : img.comp) c.dc_pred = 0;
        marker = kNoMarker;
        todo = restart_interval ? restart_interval : 0x7fffffff;
        eob_run = 0;
    }

    // baseline block (stb:2142-2193); the value stored is the coefficient BEFORE dequantisation
    bool block_baseline(int16_t* data, int n) {
        Component& c = img.comp[n];
        if (code_bits < 16) grow();
        const int t = huff_decode(huff_dc[c.hd]);
        if (t < 0 || t > 15) return fail("bad huffman code");
        memset(data, 0, 64 * sizeof(int16_t));
        const int diff = t ? extend_receive(t) : 0;
        const int dc = c.dc_pred + diff;
        c.dc_pred = dc;
        data[0] = int16_t(dc);
        int k = 1;
        do {
            const int rs = huff_decode(huff_ac[c.ha]);
            if (rs < 0) return fail("bad huffman code");
            const int s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (rs != 0xf0) break;
                k += 16;
            } else {
                k += r;
                const unsigned zig = kDezigzag[k < 79 ? k : 78];
                ++k;
                data[zig] = int16_t(extend_receive(s));
            }
        } while (k < 64);
        return true;
    }
    bool block_prog_dc(int16_t* data, int n) { // stb:2195-2219
        if (spec_end != 0) return fail("can't merge dc and ac");
        Component& c = img.comp[n];
        if (code_bits < 16) grow();
        if (succ_high == 0) {
            memset(data, 0, 64 * sizeof(int16_t));
            const int t = huff_decode(huff_dc[c.hd]);
            if (t < 0 || t > 15) return fail("can't merge dc and ac");
            const int diff = t ? extend_receive(t) : 0;
            const int dc = c.dc_pred + diff;
            c.dc_pred = dc;
            data[0] = int16_t(uint32_t(dc) << succ_low); // two's-complement shift of a possibly negative value, well defined
        } else if (get_bit()) {
            data[0] = int16_t(data[0] + int16_t(1 << succ_low));
        }
        return true;
    }
    bool block_prog_ac(int16_t* data, int n) { // stb:2223-2338
        if (spec_start == 0) return fail("can't merge dc and ac");
        const Huffman& hac = huff_ac[img.comp[n].ha];
        if (succ_high == 0) {
            const int shift = succ_low;
            if (eob_run) {
                --eob_run;
                return true;
            }
            int k = spec_start;
            do {
                const int rs = huff_decode(hac);
                if (rs < 0) return fail("bad huffman code");
                const int s = rs & 15;
                int r = rs >> 4;
                if (s == 0) {
                    if (r < 15) {
                        eob_run = 1 << r;
                        if (r) eob_run += get_bits(r);
                        --eob_run;
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    const unsigned zig = kDezigzag[k < 79 ? k : 78];
                    ++k;
                    data[zig] = int16_t(uint32_t(extend_receive(s)) << shift);
                }
            } while (k <= spec_end);
        } else {
            const int16_t bit = int16_t(1 << succ_low);
            auto refine = [&](int16_t* q) { // one correction bit for a coefficient that is already non-zero
                if (get_bit())
                    if ((*q & bit) == 0) *q = int16_t(*q > 0 ? *q + bit : *q - bit);
            };
            if (eob_run) {
                --eob_run;
                for (int k = spec_start; k <= spec_end; ++k) {
                    int16_t* q = &data[kDezigzag[k]];
                    if (*q != 0) refine(q);
                }
            } else {
                int k = spec_start;
                do {
                    const int rs = huff_decode(hac);
                    if (rs < 0) return fail("bad huffman code");
                    int s = rs & 15, r = rs >> 4;
                    if (s == 0) {
                        if (r < 15) {
                            eob_run = (1 << r) - 1;
                            if (r) eob_run += get_bits(r);
                            r = 64; // force end of block
                        }
                        // r == 15: a run of 16 zeros, handled by the loop below with s == 0
                    } else {
                        if (s != 1) return fail("bad huffman code");
                        s = get_bit() ? bit : -bit;
                    }
                    while (k <= spec_end) { // advance by r zero coefficients, refining the non-zero ones passed on the way
                        int16_t* q = &data[kDezigzag[k++]];
                        if (*q != 0) {
                            refine(q);
                        } else {
                            if (r == 0) {
                                *q = int16_t(s);
                                break;
                            }
                            --r;
                        }
                    }
                } while (k <= spec_end);
            }
        }
        return true;
    }

    bool restart_check(bool& stop) { // stb:2895-2901: count down the restart interval after every MCU
        stop = false;
        if (--todo <= 0) {
            if (code_bits < 24) grow();
            if (!(marker >= 0xd0 && marker <= 0xd7)) {
                stop = true;
                return true;
            }
            reset();
        }
        return true;
    }

    bool decode_scan() { // stb:2875-2996
        reset();
        if (scan_n == 1) { // non-interleaved: the component's own block grid, row by row
            const int n = order[0];
            Component& c = img.comp[n];
            const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
            for (int j = 0; j < h; ++j)
                for (int i = 0; i < w; ++i) {
                    int16_t* data = c.coeff.data() + 64 * size_t(i + j * c.blocks_w);
                    bool ok;
                    if (!img.progressive) ok = block_baseline(data, n);
                    else ok = spec_start == 0 ? block_prog_dc(data, n) : block_prog_ac(data, n);
                    if (!ok) return false;
                    bool stop;
                    restart_check(stop);
                    if (stop) return true;
                }
            return true;
        }
        for (int j = 0; j < mcu_y; ++j)
            for (int i = 0; i < mcu_x; ++i) {
                for (int k = 0; k < scan_n; ++k) {
                    const int n = order[k];
                    Component& c = img.comp[n];
                    for (int y = 0; y < c.v; ++y)
                        for (int x = 0; x < c.h; ++x) {
                            const int x2 = i * c.h + x, y2 = j * c.v + y;
                            int16_t* data = c.coeff.data() + 64 * size_t(x2 + y2 * c.blocks_w);
                            if (!(img.progressive ? block_prog_dc(data, n) : block_baseline(data, n))) return false;
                        }
                }
                bool stop;
                restart_check(stop);
                if (stop) return true;
            }
        return true;
    }

    bool process_marker(int m) { // stb:3025-3120
        switch (m) {
        case kNoMarker: return fail("expected marker");
        case 0xDD:
            if (get16() != 4) return fail("bad DRI len");
            restart_interval = get16();
            return true;
        case 0xDB: {
            int L = get16() - 2;
            while (L > 0) {
                const int q = get8();
                const int prec = q >> 4, t = q & 15;
                if (prec != 0 && prec != 1) return fail("bad DQT type");
                if (t > 3) return fail("bad DQT table");
                for (int i = 0; i < 64; ++i) img.dequant[t][kDezigzag[i]] = uint16_t(prec ? get16() : get8());
                L -= prec ? 129 : 65;
            }
            return L == 0 ? true : fail("bad DQT len");
        }
        case 0xC4: {
            int L = get16() - 2;
            while (L > 0) {
                int sizes[16], n = 0;
                const int q = get8();
                const int tc = q >> 4, th = q & 15;
                if (tc > 1 || th > 3) return fail("bad DHT header");
                for (int i = 0; i < 16; ++i) {
                    sizes[i] = get8();
                    n += sizes[i];
                }
                if (n > 256) return fail("bad DHT header");
                L -= 17;
                Huffman& h = tc == 0 ? huff_dc[th] : huff_ac[th];
                if (!h.build(sizes)) return fail("bad code lengths");
                for (int i = 0; i < n; ++i) h.values[i] = uint8_t(get8());
                L -= n;
            }
            return L == 0 ? true : fail("bad DHT len");
        }
        default: break;
        }
        if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE) {
            int L = get16();
            if (L < 2) return fail("bad APP/COM len");
            L -= 2;
            if (m == 0xE0 && L >= 5) {
                static const uint8_t tag[5] = {'J', 'F', 'I', 'F', 0};
                bool ok = true;
                for (int i = 0; i < 5; ++i)
                    if (get8() != tag[i]) ok = false;
                L -= 5;
                if (ok) jfif = true;
            } else if (m == 0xEE && L >= 12) {
                static const uint8_t tag[6] = {'A', 'd', 'o', 'b', 'e', 0};
                bool ok = true;
                for (int i = 0; i < 6; ++i)
                    if (get8() != tag[i]) ok = false;
                L -= 6;
                if (ok) {
                    get8();
                    get16();
                    get16();
                    app14 = get8();
                    L -= 6;
                }
            }
            skip(L);
            return true;
        }
        return fail("unknown marker");
    }

    bool frame_header() { // stb:3195-3280
        const int Lf = get16();
        if (Lf < 11) return fail("bad SOF len");
        if (get8() != 8) return fail("only 8-bit JPEG is supported");
        img.height = get16();
        img.width = get16();
        if (img.height == 0) return fail("no header height");
        if (img.width == 0) return fail("0 width");
        const int c = get8();
        if (c != 3 && c != 1 && c != 4) return fail("bad component count");
        if (c == 4) return fail("4-component (CMYK/YCCK) JPEG is not supported");
        img.n_comp = c;
        // stb:3223 (stbi__mad3sizes_valid): width * height * components must fit an int — also what keeps a damaged
        // header from sizing the coefficient arrays below
        if (uint64_t(img.width) * uint64_t(img.height) * uint64_t(c) > 0x7fffffffull) return fail("too large");
        if (Lf != 8 + 3 * c) return fail("bad SOF len");
        rgb_ids = 0;
        for (int i = 0; i < c; ++i) {
            static const uint8_t rgb[3] = {'R', 'G', 'B'};
            Component& k = img.comp[i];
            k.id = get8();
            if (c == 3 && k.id == rgb[i]) ++rgb_ids;
            const int q = get8();
            k.h = q >> 4;
            k.v = q & 15;
            if (!k.h || k.h > 4) return fail("bad H");
            if (!k.v || k.v > 4) return fail("bad V");
            k.tq = get8();
            if (k.tq > 3) return fail("bad TQ");
        }
        int h_max = 1, v_max = 1;
        for (int i = 0; i < c; ++i) {
            if (img.comp[i].h > h_max) h_max = img.comp[i].h;
            if (img.comp[i].v > v_max) v_max = img.comp[i].v;
        }
        img.h_max = h_max;
        img.v_max = v_max;
        const int mcu_w = h_max * 8, mcu_h = v_max * 8;
        mcu_x = (img.width + mcu_w - 1) / mcu_w;
        mcu_y = (img.height + mcu_h - 1) / mcu_h;
        for (int i = 0; i < c; ++i) {
            Component& k = img.comp[i];
            k.x = (img.width * k.h + h_max - 1) / h_max;
            k.y = (img.height * k.v + v_max - 1) / v_max;
            k.w2 = mcu_x * k.h * 8;
            k.h2 = mcu_y * k.v * 8;
            k.blocks_w = k.w2 / 8;
            k.blocks_h = k.h2 / 8;
            k.coeff.assign(size_t(k.blocks_w) * k.blocks_h * 64, 0);
        }
        return true;
    }

    bool scan_header() { // stb:3123-3161
        const int Ls = get16();
        scan_n = get8();
        if (scan_n < 1 || scan_n > 4 || scan_n > img.n_comp) return fail("bad SOS component count");
        if (Ls != 6 + 2 * scan_n) return fail("bad SOS len");
        for (int i = 0; i < scan_n; ++i) {
            const int id = get8(), q = get8();
            int which = 0;
            for (; which < img.n_comp; ++which)
                if (img.comp[which].id == id) break;
            if (which == img.n_comp) return fail("bad SOS component");
            img.comp[which].hd = q >> 4;
            img.comp[which].ha = q & 15;
            if (img.comp[which].hd > 3 || img.comp[which].ha > 3) return fail("bad huffman table index");
            order[i] = which;
        }
        spec_start = get8();
        spec_end = get8();
        const int aa = get8();
        succ_high = aa >> 4;
        succ_low = aa & 15;
        if (img.progressive) {
            if (spec_start > 63 || spec_end > 63 || spec_start > spec_end || succ_high > 13 || succ_low > 13) return fail("bad SOS");
        } else {
            if (spec_start != 0 || succ_high != 0 || succ_low != 0) return fail("bad SOS");
            spec_end = 63;
        }
        // the tables this scan decodes with must exist by now (stb decodes against whatever its table memory holds;
        // here that would be undefined behaviour on untrusted input).  Refinement scans of DC coefficients read raw bits.
        for (int i = 0; i < scan_n; ++i) {
            const auto& c = img.comp[order[i]];
            const bool needs_dc = img.progressive ? (spec_start == 0 && succ_high == 0) : true;
            const bool needs_ac = img.progressive ? spec_start != 0 : true;
            if ((needs_dc && !huff_dc[c.hd].defined) || (needs_ac && !huff_ac[c.ha].defined)) return fail("undefined huffman table");
        }
        return true;
    }

    bool run() { // stb:3283-3345
        int m = get_marker();
        if (m != 0xd8) return fail("no SOI");
        m = get_marker();
        while (!(m == 0xc0 || m == 0xc1 || m == 0xc2)) {
            if (!process_marker(m)) return false;
            m = get_marker();
            while (m == kNoMarker) {
                if (at_eof()) return fail("no SOF");
                m = get_marker();
            }
        }
        img.progressive = m == 0xc2;
        if (!frame_header()) return false;
        m = get_marker();
        while (m != 0xd9) {
            if (m == 0xda) {
                if (!scan_header()) return false;
                if (!decode_scan()) return false;
                if (marker == kNoMarker) { // trailing zeros after the entropy-coded data
                    while (!at_eof()) {
                        const int x = get8();
                        if (x == 255) {
                            marker = uint8_t(get8());
                            break;
                        }
                    }
                }
            } else if (m == 0xdc) {
                const int Ld = get16(), NL = get16();
                if (Ld != 4) return fail("bad DNL len");
                if (NL != img.height) return fail("bad DNL height");
            } else {
                if (!process_marker(m)) return false;
            }
            if (at_eof() && marker == kNoMarker) return fail("no EOI");
            m = get_marker();
        }
        img.is_rgb = img.n_comp == 3 && (rgb_ids == 3 || (app14 == 0 && !jfif));
        return true;
    }
};

} // namespace

bool decode_coefficients(const uint8_t* data, size_t n_bytes, CoefficientImage& out) {
    out = CoefficientImage();
    if (!data || n_bytes < 4) {
        out.error = "not a JPEG file";
        return false;
    }
    try { // the coefficient arrays of a (legitimately) huge frame may not fit: an error, never an exception for the C ABI above
        Decoder d(data, n_bytes, out);
        return d.run();
    } catch (const std::exception&) {
        out.error = "out of memory";
        return false;
    }
}

} // namespace rtj
