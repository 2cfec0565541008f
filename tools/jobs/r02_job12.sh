#!/bin/bash
# Round-2 GPU job 12: C4 — occupancy variants of the quantised-node kernel, then an ncu capture of it
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python tools/c4_ab_lib.py full q3 q4 q3@RT_PT_REFILL=8 q3@RT_PT_REFILL=24 q4@RT_PT_REFILL=8 > gpurun_out/c4_ab2.log 2>&1; cat gpurun_out/c4_ab2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_step_pt -s 2 -c 1 -o gpurun_out/r02_prof_c4_quant -f python tools/c4_small.py > gpurun_out/ncu_c4_quant.log 2>&1; tail -n 2 gpurun_out/ncu_c4_quant.log
