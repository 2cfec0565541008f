// rt_jpeg.cuh — device JPEG writer (rt_jpeg.cu): the reference's stbi_write_jpg output stage as CUDA kernels.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace rtd {

struct JpegState; // tables of the current quality + scratch buffers, one per context
JpegState* jpeg_create();
void jpeg_destroy(JpegState* s);
// upper bound of the file size for a w x h image (any quality)
size_t jpeg_max_bytes(int w, int h);
// rgb8_dev: h rows of w RGB bytes, first row = top of the picture (what stbi_write_jpg receives, main.cu:491).
// Writes the finished file to out_host (capacity cap); out_host == nullptr only reports the size.
// Synchronises `st`.  ms_device: device time of the encoder passes (CUDA events), may be nullptr.
cudaError_t jpeg_encode(JpegState* s, const uint8_t* rgb8_dev, int w, int h, int quality, uint8_t* out_host, size_t cap,
                        size_t* n_bytes, cudaStream_t st, float* ms_device);

} // namespace rtd
