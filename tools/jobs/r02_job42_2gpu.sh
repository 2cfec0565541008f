#!/bin/bash
# Round-2 GPU job 42 (gpurun --gpus 2): ncu --set full of k_reduce_tonemap at 8K with two members in ONE process (rt_multi_*, events
# between the devices, no spinning kernel), for the NVLink side of the framebuffer passes; and the event-timed reduce of the same frame
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 ./apps/render_scene --scene book1_final --width 7680 --height 4320 --spp 8 --gpus 2 --out gpurun_out/m2.ppm > gpurun_out/multi2_8k.log 2>&1; cat gpurun_out/multi2_8k.log; rm -f gpurun_out/m2.ppm
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reduce_tonemap -c 2 -o gpurun_out/r02f_prof_reduce_8k_2gpu -f ./apps/render_scene --scene book1_final --width 7680 --height 4320 --spp 8 --gpus 2 --out gpurun_out/m2.ppm > gpurun_out/ncu_reduce_2gpu.log 2>&1; tail -n 3 gpurun_out/ncu_reduce_2gpu.log; rm -f gpurun_out/m2.ppm
