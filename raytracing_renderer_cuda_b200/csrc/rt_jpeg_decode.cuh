// rt_jpeg_decode.cuh — device half of the JPEG reader (rt_jpeg_decode.cu): coefficients -> pixels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rt_jpeg_decode_host.hpp"

namespace rtd {

// Uploads the coefficient planes and runs dequantisation + IDCT, up-sampling and colour conversion.
// out_dev: width * height * channels floats (channels = 3, or 1 for a one-component file), value = byte / 255.f,
// row 0 = top of the picture: what stbi_loadf returns (main.cu:376-380).  Synchronises `st`.
cudaError_t jpeg_pixels_device(const rtj::CoefficientImage& img, float* out_dev, cudaStream_t st, float* ms_device);

} // namespace rtd
