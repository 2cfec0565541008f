#!/bin/bash
# Round-2 GPU job 30: ncu --set full of k_wf_step_warp on C3 (perlin_motion)
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_step_warp -s 3 -c 1 -o gpurun_out/r02f_prof_c3_warp -f python tools/c2_once.py perlin_motion > gpurun_out/r02f_ncu_c3.log 2>&1; tail -n 2 gpurun_out/r02f_ncu_c3.log
