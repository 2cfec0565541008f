#!/bin/bash
# Round-2 GPU job 16: pipelined e2e of the new bench.py, framebuffer passes at 8K (timing + ncu), 9/10 CTAs per SM for the C1 kernel
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline --no-other-configs > gpurun_out/r02_bench_e2e.json 2> gpurun_out/r02_bench_e2e.err; tail -n 5 gpurun_out/r02_bench_e2e.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_e2e.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['serial'])"
AB_NO_MEGA=1 AB_CASES=c1 timeout 600 python tools/ab_test.py base mb9 mb10 base > gpurun_out/ab_mb9.log 2>&1; cat gpurun_out/ab_mb9.log
timeout 300 python tools/tonemap_8k_once.py > gpurun_out/tonemap_8k.log 2>&1; cat gpurun_out/tonemap_8k.log
timeout 600 ncu --set full --clock-control none -k regex:"k_tonemap|k_reduce_tonemap" -s 6 -c 2 -o gpurun_out/r02_prof_tonemap_8k -f python tools/tonemap_8k_once.py > gpurun_out/ncu_tonemap.log 2>&1; tail -n 2 gpurun_out/ncu_tonemap.log
