#!/bin/bash
# Round-2 GPU job 24: C4, final build: leaf-step and refill thresholds of k_wf_step_pt at 28 warps/SM; armed k_wf_tail launch under ncu
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python tools/c4_ab_lib.py full cur cur@RT_PT_LEAF_LANES=4 cur@RT_PT_LEAF_LANES=8 cur@RT_PT_REFILL=12 cur@RT_PT_REFILL=20 cur@RT_NO_L2_PERSIST=1 cur > gpurun_out/c4_ab_knobs2.log 2>&1; cat gpurun_out/c4_ab_knobs2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_tail -c 4 -o gpurun_out/r02f_prof_c1_tail -f python tools/c1_once.py > gpurun_out/r02f_ncu_c1t.log 2>&1; tail -n 1 gpurun_out/r02f_ncu_c1t.log
