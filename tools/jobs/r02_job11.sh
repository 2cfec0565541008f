#!/bin/bash
# Round-2 GPU job 11: quantised 4-wide nodes — parity (trace vs reference kernel / brute force, image equality), then the C4 A/B
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_c4_parity.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "c4 or closest or bvh_equals or million or trace_matches" > gpurun_out/pytest_q.log 2>&1; tail -n 15 gpurun_out/pytest_q.log | cut -c1-250
timeout 900 python tools/c4_ab.py full > gpurun_out/c4_ab.log 2>&1; cat gpurun_out/c4_ab.log
