#!/bin/bash
# Round-2 GPU job 4 (1 GPU): GPU suite, smoke, the new bench.py (other_configs, clean reference arm).
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -n 8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_c1.json 2> gpurun_out/r02_bench_c1.err; tail -n 5 gpurun_out/r02_bench_c1.err; cut -c1-600 gpurun_out/r02_bench_c1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_c1.json 2> gpurun_out/r02_bench_ref_c1.err; cut -c1-400 gpurun_out/r02_bench_ref_c1.json
python tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; tail -n 4 gpurun_out/e2e_probe.log
