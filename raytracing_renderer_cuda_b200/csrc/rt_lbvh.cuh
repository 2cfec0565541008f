// rt_lbvh.cuh — GPU LBVH build (Morton codes over the motion-union AABBs -> radix sort ->
// Karras topology -> bottom-up refit) into the flattened 64-byte node array.
#pragma once

#include <cuda_runtime.h>

#include "rt_device.cuh"

namespace rtd {

// boxes: n x 6 floats (lo.xyz, hi.xyz) on the device, prim index = array index.
// nodes: n-1 BvhNode, root = 0.  n >= 2.
cudaError_t lbvh_build(const float* boxes_dev, uint32_t n, BvhNode* nodes_dev, cudaStream_t st, float* ms,
                       uint32_t* depth);

} // namespace rtd
