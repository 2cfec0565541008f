#!/bin/bash
# Round-2 GPU job 29: binary BVH nodes in centre / half-extent form: full GPU suite (parity), then A/B against the min/max form
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu_29.log 2>&1; tail -n 12 gpurun_out/pytest_gpu_29.log | cut -c1-300
timeout 900 python tools/ab_test.py cur ch cur ch > gpurun_out/ab_ch.log 2>&1; cat gpurun_out/ab_ch.log
