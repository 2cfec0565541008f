#!/bin/bash
# Builds gpurun_variants/librt_ring_{e,r,s,b4,all}.so for tools/ring_ab.py: only rt_wavefront.cu differs, the other objects
# come from build/ (run `make lib` first).  Flags: WF_RING_EARLY_CLAIM, WF_RING_RELAXED_PUBLISH (unsafe, A/B only),
# WF_RING_ACC_STREAM (unsafe, A/B only) — see profiles/r01_ring.md; WF_RING_BATCH=4 (batch claims: untested hypothesis
# that the counters of the fullest class are the bulk's bottleneck).
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
build_v() {
    name=$1; shift
    $NV "$@" -c raytracing_renderer_cuda_b200/csrc/rt_wavefront.cu -o gpurun_variants/wf_$name.o
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_variants/librt_$name.so gpurun_variants/wf_$name.o \
        build/rt_api.o build/rt_kernels.o build/rt_lbvh.o build/rt_jpeg.o build/rt_jpeg_decode.o build/rt_host.o build/rt_bvh_host.o \
        build/rt_jpeg_decode_host.o -lcudart
    rm gpurun_variants/wf_$name.o
}
build_v ring_e -DWF_RING_EARLY_CLAIM &
build_v ring_r -DWF_RING_RELAXED_PUBLISH &
build_v ring_s -DWF_RING_ACC_STREAM &
build_v ring_b4 -DWF_RING_BATCH=4u &
build_v ring_all -DWF_RING_EARLY_CLAIM -DWF_RING_RELAXED_PUBLISH -DWF_RING_ACC_STREAM &
wait
ls -la gpurun_variants/
