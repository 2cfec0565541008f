// rt_internal.hpp — the opaque handles of include/rt_api.h, shared by the translation units that implement the C-ABI
// (rt_api.cu: contexts, scenes, render entries; rt_multi.cu: the multi-GPU entries).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "../../include/rt_api.h"
#include "rt_jpeg.cuh"
#include "rt_kernels.cuh"

namespace rtd {
void set_error_message(const char* msg); // rt_api.cu: the message rt_last_error() returns on this thread
}

struct rt_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int sm_count = 0;             // multiProcessorCount of the device (launch sizing)
    cudaMemPool_t pool = nullptr; // scene allocations (stream-ordered, private to the context)
    unsigned long long* d_ray_counter = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float4* accum = nullptr; // scratch for rt_render / rt_render_accum
    size_t accum_px = 0;
    float* out_rgb = nullptr;
    size_t out_px = 0;
    rtd::WavefrontState* wf = nullptr;
    rtd::JpegState* jpg = nullptr; // device JPEG writer (rt_jpeg.cu), created on first use
    uint8_t* rgb8 = nullptr;       // flipped, quantised frame: input of the JPEG writer
    size_t rgb8_px = 0;
    // cudaArray allocations cost milliseconds; arrays of destroyed scenes are kept for the next scene of the
    // same image size (a frame loop that re-uploads its scene every frame then allocates nothing)
    struct CachedArray {
        int32_t width, height;
        cudaArray_t arr;
    };
    std::vector<CachedArray> array_cache;
    std::mutex cache_mutex; // a scene may be destroyed on one host thread while another creates the next one (frame pipelines)
};

struct rt_scene {
    rt_context* ctx = nullptr;
    rtd::DScene d{};
    rt_scene_info info{};
    std::vector<void*> allocs;
    std::vector<rt_context::CachedArray> arrays;
    std::vector<cudaTextureObject_t> texobjs;
};

