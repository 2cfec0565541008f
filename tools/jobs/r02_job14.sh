#!/bin/bash
# Round-2 GPU job 14: in-place completion of the thin end of a frame (wf_tail_mode): parity with it on, then the threshold sweep
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
RT_WF_TAIL_PATHS=150000 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_emitter_sampling.py -m gpu -q --timeout 600 -x -k "not million and not bvh_equals and not trace_ and not shading_step and not round_toward and not 8k" > gpurun_out/pytest_tail.log 2>&1; tail -n 8 gpurun_out/pytest_tail.log | cut -c1-200
AB_NO_MEGA=1 timeout 900 python tools/ab_test.py listc tail tail@RT_WF_TAIL_PATHS=16384 tail@RT_WF_TAIL_PATHS=65536 tail@RT_WF_TAIL_PATHS=150000 tail@RT_WF_TAIL_PATHS=300000 tail@RT_WF_TAIL_PATHS=1000000 tail > gpurun_out/ab_tail.log 2>&1; cat gpurun_out/ab_tail.log
