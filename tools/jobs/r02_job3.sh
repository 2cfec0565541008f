#!/bin/bash
# Round-2 GPU job 3: whole GPU suite (no -x), smoke, C1 bench + ncu of the default (rolled constant-bank list) build.
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -n 30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02_diet_c1.json 2> gpurun_out/r02_diet_c1.err; cut -c1-300 gpurun_out/r02_diet_c1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_step -s 4 -c 1 -o gpurun_out/r02_prof_c1_diet -f python tools/c1_once.py > gpurun_out/ncu_c1_diet.log 2>&1; tail -n 2 gpurun_out/ncu_c1_diet.log
