#!/bin/bash
# Round-2 GPU job 18: k_wf_tail (parity + threshold sweep), sign-selected / PRMT node visit of k_wf_step_pt (parity + C4 A/B, with and without the L1 prefetch)
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu_18.log 2>&1; tail -n 15 gpurun_out/pytest_gpu_18.log | cut -c1-300
AB_NO_MEGA=1 timeout 900 python tools/ab_test.py base cur@RT_WF_TAIL_PATHS=0 cur@RT_WF_TAIL_PATHS=16384 cur cur@RT_WF_TAIL_PATHS=150000 cur@RT_WF_TAIL_PATHS=300000 base > gpurun_out/ab_tailk.log 2>&1; cat gpurun_out/ab_tailk.log
timeout 900 python tools/c4_ab_lib.py full base cur pf > gpurun_out/c4_ab_travq.log 2>&1; cat gpurun_out/c4_ab_travq.log
