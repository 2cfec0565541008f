// oracle/ref_stb.cpp — TEST INFRASTRUCTURE.  The reference's own JPEG writer: its vendored stb_image_write.h is
// compiled from where it lies under /root/reference (nothing is copied), exactly as src/main.cu:11-12 includes it,
// and exposed through one C function so the tests can compare bytes.  Built into oracle/_ref/libref_stb.so.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "libs/stb/stb_image_write.h"

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
