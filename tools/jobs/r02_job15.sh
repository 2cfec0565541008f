#!/bin/bash
# Round-2 GPU job 15: compiler / occupancy knobs of the C1 kernel on the shipped source
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
AB_NO_MEGA=1 timeout 900 python tools/ab_test.py base mb7 mb6 exp base > gpurun_out/ab_knobs.log 2>&1; cat gpurun_out/ab_knobs.log
