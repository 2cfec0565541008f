#!/bin/bash
# Round-2 GPU job 22: pipelined vs serial e2e of bench.py with and without k_wf_tail (two runs each, one box)
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for rep in 1 2; do
for tp in default 0; do
  if [ $tp = default ]; then unset RT_WF_TAIL_PATHS; else export RT_WF_TAIL_PATHS=$tp; fi
  timeout 600 python bench.py --no-cpu-baseline --no-other-configs > gpurun_out/e2e_probe_$tp.json 2> gpurun_out/e2e_probe_$tp.err
  python -c "
import json; d=json.load(open('gpurun_out/e2e_probe_$tp.json')); print('tail=$tp', 'device', round(d['ms_per_step'],3), 'pipelined', round(d['e2e']['ms_per_step'],3), 'serial', round(d['e2e']['serial']['ms_per_step'],3), 'jpeg', round(d['output_stage']['ms_per_step'],3))"
done; done
