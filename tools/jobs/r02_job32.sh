#!/bin/bash
# Round-2 GPU job 32: the persistent-lane kernel (quantised 4-wide nodes) on the small BVH scenes C2 / C3, against the warp-chunk kernel
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
AB_NO_MEGA=1 AB_CASES=c2,c3 timeout 900 python tools/ab_test.py cur cur@RT_WF_GRAIN=pt cur@RT_WF_GRAIN=pt,RT_PT_REFILL=8 cur@RT_WF_GRAIN=pt,RT_PT_REFILL=24 cur@RT_WF_GRAIN=cta cur > gpurun_out/ab_grain_small.log 2>&1; cat gpurun_out/ab_grain_small.log
