#!/bin/bash
# Round-2 GPU job 1 (one B200, through gpurun): reference-kernel golden for the C4 generator at n = 10^4, baseline
# bench lines of the round-1 kernels on this box, and ncu --set full captures of the two kernels the round works on.
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python tools/make_gpu_golden_c4.py > gpurun_out/golden_c4.log 2>&1; tail -n 3 gpurun_out/golden_c4.log
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02_base_c1.json 2> gpurun_out/r02_base_c1.err
for c in c2 c3 c4; do python bench.py --config $c --steps 3 --no-cpu-baseline > gpurun_out/r02_base_$c.json 2> gpurun_out/r02_base_$c.err; done
cut -c1-400 gpurun_out/r02_base_c*.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_step -s 4 -c 1 -o gpurun_out/r02_prof_c1_base -f python tools/c1_once.py > gpurun_out/ncu_c1_base.log 2>&1; tail -n 2 gpurun_out/ncu_c1_base.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_step_pt -s 2 -c 1 -o gpurun_out/r02_prof_c4_base -f python tools/c4_small.py > gpurun_out/ncu_c4_base.log 2>&1; tail -n 2 gpurun_out/ncu_c4_base.log
ls -la gpurun_out
