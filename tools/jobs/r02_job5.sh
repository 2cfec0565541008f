#!/bin/bash
# Round-2 GPU job 5: (a) why did the C1 frame take 7.44 ms in job 3 and 8.15 ms in job 4 with an identical kernel?  A/B of the
# two libraries on ONE box.  (b) first run of the record-block kernel (RT_WF_GRAIN=blk): parity subset, then A/B.
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
AB_NO_MEGA=1 timeout 600 python tools/ab_test.py r01 rolled j4 cur rolled j4 > gpurun_out/ab_place.log 2>&1; cat gpurun_out/ab_place.log
RT_WF_GRAIN=blk RT_PARITY_LOG=gpurun_out/parity_blk.jsonl timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_emitter_sampling.py -m gpu -q --timeout 600 -x -k "not million and not bvh_equals and not trace_ and not shading_step" > gpurun_out/pytest_blk.log 2>&1; tail -n 12 gpurun_out/pytest_blk.log
AB_NO_MEGA=1 timeout 600 python tools/ab_test.py cur cur@RT_WF_GRAIN=blk cur cur@RT_WF_GRAIN=blk > gpurun_out/ab_blk.log 2>&1; cat gpurun_out/ab_blk.log
