#!/bin/bash
# Round-2 GPU job 2: the GPU suite on the instruction-diet build (new Perlin walk, constant-bank sphere list, FastDiv path
# mapping, one-thread chunk location), smoke, A/B frame times against the round-1 library on the SAME box, ncu capture.
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
AB_NO_MEGA=1 timeout 600 python tools/ab_test.py r01 main rolled nolist r01 main > gpurun_out/ab_diet.log 2>&1; cat gpurun_out/ab_diet.log
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02_diet_c1.json 2> gpurun_out/r02_diet_c1.err; cut -c1-300 gpurun_out/r02_diet_c1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_step -s 4 -c 1 -o gpurun_out/r02_prof_c1_diet -f python tools/c1_once.py > gpurun_out/ncu_c1_diet.log 2>&1; tail -n 2 gpurun_out/ncu_c1_diet.log
