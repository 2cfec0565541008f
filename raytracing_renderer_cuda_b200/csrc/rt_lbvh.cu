// rt_lbvh.cu — GPU LBVH build (placeholder until the Karras builder lands in this file).
#include "rt_lbvh.cuh"

namespace rtd {

cudaError_t lbvh_build(const float*, uint32_t, BvhNode*, cudaStream_t, float*, uint32_t*) { return cudaErrorNotSupported; }

} // namespace rtd
