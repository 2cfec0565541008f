#!/bin/bash
# Round-2 GPU job 19: C4 A/B of child ordering (full sort / nearest first / none) and of 28 warps per SM (128-thread CTAs, 72 registers)
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python tools/c4_ab_lib.py full cur nosort near1 t128 t128b cur > gpurun_out/c4_ab_order.log 2>&1; cat gpurun_out/c4_ab_order.log
