#!/bin/bash
# Round-2 GPU job 41: C4 A/B — postponed leaves in k_wf_step_pt (tested when 8 / 16 / 24 lanes hold one), then parity of the 16-lane variant
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python tools/c4_ab_lib.py full cur pp8 pp16 pp24 cur > gpurun_out/c4_ab_postpone.log 2>&1; cat gpurun_out/c4_ab_postpone.log
RT_B200_LIB=$PWD/gpurun_variants/librt_pp16.so timeout 600 python -m pytest tests/test_gpu_c4_parity.py tests/test_gpu_parity.py -m gpu -q -k "c4 or million or bvh_equals" > gpurun_out/pytest_pp16.log 2>&1; tail -n 3 gpurun_out/pytest_pp16.log
