#!/bin/bash
# Round-2 GPU job 28: CTA size of k_wf_tail (fewer, larger CTAs give denser class segments): 128 x 8, 256 x 4, 512 x 2 per SM
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
AB_NO_MEGA=1 timeout 900 python tools/ab_test.py cur t256 t512 t256@RT_WF_TAIL_PATHS=262144 t512@RT_WF_TAIL_PATHS=262144 cur > gpurun_out/ab_tail_cta_size.log 2>&1; cat gpurun_out/ab_tail_cta_size.log
