#!/bin/bash
# Round-2 GPU job 25: k_wf_tail as a CTA-local wavefront: parity, then the threshold sweep against the one-thread-per-path form (cur)
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tail.py tests/test_gpu_parity.py tests/test_gpu_emitter_sampling.py -m gpu -x -q --timeout 600 > gpurun_out/pytest_tail2.log 2>&1; tail -n 12 gpurun_out/pytest_tail2.log | cut -c1-300
AB_NO_MEGA=1 timeout 900 python tools/ab_test.py cur tail2 tail2@RT_WF_TAIL_PATHS=32768 tail2@RT_WF_TAIL_PATHS=262144 tail2@RT_WF_TAIL_PATHS=524288 tail2@RT_WF_TAIL_PATHS=1200000 cur > gpurun_out/ab_tail2.log 2>&1; cat gpurun_out/ab_tail2.log
