#!/bin/bash
# Round-2 GPU job 34: the randomized differential parity tests and the tail tests on the final build
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
RT_PARITY_LOG=gpurun_out/parity_random.jsonl timeout 900 python -m pytest tests/test_gpu_random_scenes.py tests/test_gpu_tail.py -m gpu -q --timeout 600 > gpurun_out/pytest_random.log 2>&1; tail -n 30 gpurun_out/pytest_random.log | cut -c1-250
cat gpurun_out/parity_random.jsonl | cut -c1-300
