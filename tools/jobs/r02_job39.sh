#!/bin/bash
# Round-2 GPU job 39: hand-written scan / radix sort instead of CUB (LBVH build, 4-wide compaction, JPEG writer): complete GPU suite,
# then the build and output-stage timings (C4 e2e = scene upload + LBVH + collapse + quantisation + frame; JPEG probe)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_39.log 2>&1; tail -n 4 gpurun_out/pytest_gpu_39.log
timeout 600 python bench.py --config c4 --steps 2 --no-cpu-baseline > gpurun_out/bench_c4_39.json 2> gpurun_out/bench_c4_39.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_39.json')); print('c4', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
timeout 300 python tools/jpeg_probe.py > gpurun_out/jpeg_probe_39.log 2>&1; tail -n 12 gpurun_out/jpeg_probe_39.log
