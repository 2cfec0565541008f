#!/bin/bash
# Round-2 GPU job 13: split shade/extend kernels for large BVHs — image equality, then the C4 A/B (refill sweep, occupancy)
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "million" > gpurun_out/pytest_split.log 2>&1; tail -n 12 gpurun_out/pytest_split.log | cut -c1-250
timeout 1200 python tools/c4_ab_lib.py full sp sp@RT_WF_GRAIN=split sp@RT_WF_GRAIN=split,RT_EXT_REFILL=1 sp@RT_WF_GRAIN=split,RT_EXT_REFILL=8 sp@RT_WF_GRAIN=split,RT_EXT_REFILL=12 sp5@RT_WF_GRAIN=split sp@RT_WF_GRAIN=split,RT_BVH4=f > gpurun_out/c4_ab3.log 2>&1; cat gpurun_out/c4_ab3.log
