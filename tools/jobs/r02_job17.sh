#!/bin/bash
# Round-2 GPU job 17: full GPU suite on the vectorised framebuffer passes, their 8K timing + ncu, per-iteration launch list of a full C4 frame
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu_17.log 2>&1; tail -n 5 gpurun_out/pytest_gpu_17.log | cut -c1-300
timeout 300 python tools/tonemap_8k_once.py > gpurun_out/tonemap_8k_v4.log 2>&1; cat gpurun_out/tonemap_8k_v4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tonemap|k_reduce_tonemap" -s 6 -c 2 -o gpurun_out/r02_prof_tonemap_8k_v4 -f python tools/tonemap_8k_once.py > gpurun_out/ncu_tonemap_v4.log 2>&1; tail -n 2 gpurun_out/ncu_tonemap_v4.log
timeout 300 python tools/c4_full_once.py > gpurun_out/c4_full_once.log 2>&1; cat gpurun_out/c4_full_once.log
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none -k regex:k_wf_step -c 400 --csv --log-file gpurun_out/r02_launches_c4_full.csv python tools/c4_full_once.py > gpurun_out/ncu_c4_list.log 2>&1; tail -n 2 gpurun_out/ncu_c4_list.log
