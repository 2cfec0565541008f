#!/bin/bash
# Round-2 GPU job 36: two rounds of 32 entries per warp chunk (one push per 64 entries) in k_wf_step_warp, C2 / C3
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
AB_NO_MEGA=1 AB_CASES=c2,c3 timeout 900 python tools/ab_test.py cur rounds2 cur rounds2 > gpurun_out/ab_rounds2.log 2>&1; cat gpurun_out/ab_rounds2.log
