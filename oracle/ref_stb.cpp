// oracle/ref_stb.cpp — TEST INFRASTRUCTURE.  The reference's own JPEG writer: its vendored stb_image_write.h is
// compiled from where it lies under /root/reference (nothing is copied), exactly as src/main.cu:11-12 includes it,
// and exposed through one C function so the tests can compare bytes.  Built into oracle/_ref/libref_stb.so.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "libs/stb/stb_image_write.h"
#define STB_IMAGE_IMPLEMENTATION
#include "libs/stb/stb_image.h"

static void sink(void* ctx, void* data, int size) {
    auto* v = static_cast<std::vector<uint8_t>*>(ctx);
    v->insert(v->end(), static_cast<uint8_t*>(data), static_cast<uint8_t*>(data) + size);
}

// stbi_write_jpg(filename, w, h, 3, data, quality) (main.cu:491) into memory; returns the size, 0 if cap is too small
extern "C" size_t ref_stb_write_jpg(const uint8_t* rgb, int w, int h, int quality, uint8_t* out, size_t cap) {
    std::vector<uint8_t> v;
    if (!stbi_write_jpg_to_func(sink, &v, w, h, 3, rgb, quality)) return 0;
    if (v.size() > cap) return 0;
    memcpy(out, v.data(), v.size());
    return v.size();
}

// stbi_load_from_memory(file, n, &w, &h, &ch, 0): the 8-bit decode underneath stbi_loadf (main.cu:376-380; the float image
// is byte / 255.f with the gamma = scale = 1 the reference sets).  Returns channels (0 on failure); out must hold w*h*ch.
extern "C" int ref_stb_load_jpg(const uint8_t* file, int n, uint8_t* out, size_t cap, int* w, int* h) {
    int ch = 0;
    stbi_uc* px = stbi_load_from_memory(file, n, w, h, &ch, 0);
    if (!px) return 0;
    const size_t bytes = size_t(*w) * size_t(*h) * size_t(ch);
    if (bytes <= cap) memcpy(out, px, bytes);
    stbi_image_free(px);
    return bytes <= cap ? ch : 0;
}
// and the float path itself, exactly as the reference calls it
extern "C" int ref_stb_loadf_jpg(const uint8_t* file, int n, float* out, size_t cap_floats, int* w, int* h) {
    int ch = 0;
    stbi_ldr_to_hdr_scale(1.0f);
    stbi_ldr_to_hdr_gamma(1.0f);
    float* px = stbi_loadf_from_memory(file, n, w, h, &ch, 0);
    if (!px) return 0;
    const size_t cnt = size_t(*w) * size_t(*h) * size_t(ch);
    if (cnt <= cap_floats) memcpy(out, px, cnt * sizeof(float));
    stbi_image_free(px);
    return cnt <= cap_floats ? ch : 0;
}
