#!/bin/bash
# Round-2 GPU job 7: exactness of div_rz / sqrt_rz, parity suite with them, A/B (out-of-line vs inline V3 division, block kernel)
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c4_parity.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_rz.log 2>&1; tail -n 6 gpurun_out/pytest_rz.log
AB_NO_MEGA=1 timeout 600 python tools/ab_test.py j4 cur div3inl cur@RT_WF_GRAIN=blk j4 cur div3inl > gpurun_out/ab_rz.log 2>&1; cat gpurun_out/ab_rz.log
