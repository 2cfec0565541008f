#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc flags...>: builds gpurun_variants/librt_<name>.so for A/B timing on the GPU box
set -e
name=$1; shift
cd "$(dirname "$0")/.."
out=gpurun_variants/obj_$name; mkdir -p $out
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v"
for f in rt_api rt_kernels rt_wavefront rt_lbvh rt_jpeg rt_jpeg_decode rt_multi; do
  $NV "$@" -c raytracing_renderer_cuda_b200/csrc/$f.cu -o $out/$f.o 2> $out/$f.log &
done
g++ -std=c++17 -O2 -fPIC -c raytracing_renderer_cuda_b200/csrc/rt_host.cpp -o $out/rt_host.o &
g++ -std=c++17 -O2 -fPIC -c raytracing_renderer_cuda_b200/csrc/rt_bvh_host.cpp -o $out/rt_bvh_host.o &
g++ -std=c++17 -O2 -fPIC -c raytracing_renderer_cuda_b200/csrc/rt_jpeg_decode_host.cpp -o $out/rt_jpeg_decode_host.o &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_variants/librt_$name.so $out/*.o -lcudart
grep -A1 "k_wf_step" $out/rt_wavefront.log | grep -E "Used|spill" | head -4
