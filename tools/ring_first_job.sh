#!/bin/bash
# GPU job for the experimental barrier-free kernel (RT_WF_GRAIN=ring; run through gpurun, ~6 GPU-minutes): the probe
# (same frame as the default kernels, A/B frame times — passed in round 1, profiles/r01_ring.md), then the GPU parity
# tests under it, the C1 bench with and without it, and the ncu capture that round 1 had no budget for.
# Every step has its own timeout: a protocol bug would show as a hang, which must not cost a gpurun strike.
set -x
cd "$(dirname "$0")/.."
timeout 300 python tools/ring_probe.py > gpurun_out/ring_probe.log 2>&1; echo "ring_probe rc=$?"; tail -n 12 gpurun_out/ring_probe.log
if grep -q "ring probe ok" gpurun_out/ring_probe.log; then
    RT_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental_ring.py -m gpu -x -q > gpurun_out/pytest_ring_modes.log 2>&1; tail -n 3 gpurun_out/pytest_ring_modes.log
    RT_WF_GRAIN=ring timeout 300 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_ring.log 2>&1; tail -n 3 gpurun_out/pytest_ring.log
    RT_WF_GRAIN=ring timeout 120 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_c1_ring.json 2> gpurun_out/bench_c1_ring.err
    timeout 120 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_c1_default.json 2> gpurun_out/bench_c1_default.err
    cut -c1-200 gpurun_out/bench_c1_ring.json gpurun_out/bench_c1_default.json
    RT_WF_GRAIN=ring timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_wf_ring -c 1 -o gpurun_out/prof_wf_ring_c1 -f python tools/c1_once.py > gpurun_out/ncu_ring.log 2>&1
fi
