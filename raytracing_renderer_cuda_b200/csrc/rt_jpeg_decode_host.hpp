// rt_jpeg_decode_host.hpp — host half of the JPEG reader (image-texture ingest).
//
// The reference loads its texture with the vendored stb_image v2.26: stbi_loadf("textures/earth.jpg", &w, &h, &ch, 0)
// (main.cu:376-380).  A JPEG's entropy-coded segments are one serial bit stream, so the marker parsing and the Huffman
// decoding (baseline and progressive: stb_image.h:1942-2338, 2845-3345) stay on the host; they produce the quantised
// DCT coefficients of every component.  Everything after that is data-parallel and runs on the GPU (rt_jpeg_decode.cu):
// dequantisation, stb's integer IDCT, its chroma up-sampling filters, the fixed-point YCbCr -> RGB and byte/255.f.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace rtj {

struct Component {
    int id = 0, h = 1, v = 1, tq = 0; // frame header (stb_image.h:3215-3225)
    int hd = 0, ha = 0, dc_pred = 0;  // scan state
    int x = 0, y = 0;                 // effective pixels of this component (stb_image.h:3249-3250)
    int w2 = 0, h2 = 0;               // plane size padded to whole interleaved MCUs (stb_image.h:3258-3259)
    int blocks_w = 0, blocks_h = 0;   // w2 / 8, h2 / 8
    std::vector<int16_t> coeff;       // blocks_w * blocks_h * 64, natural (row-major) order inside a block, NOT dequantised
};

struct CoefficientImage {
    int width = 0, height = 0, n_comp = 0;
    int h_max = 1, v_max = 1;
    bool progressive = false;
    bool is_rgb = false; // components are R, G, B (ids 'R','G','B' or Adobe transform 0 without JFIF): no colour transform
    Component comp[4];
    uint16_t dequant[4][64] = {}; // natural order (stb_image.h:3045-3046)
    std::string error;
};

// Parses the file and decodes every scan.  Returns false (with `error`) on anything stb_image rejects or this reader does not
// cover (12-bit, arithmetic coding, CMYK).
bool decode_coefficients(const uint8_t* data, size_t n_bytes, CoefficientImage& out);

} // namespace rtj
