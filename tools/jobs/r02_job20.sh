#!/bin/bash
# Round-2 GPU job 20: C4 A/B of the persistent-lane kernel with cold lane state parked in shared memory, at 24 / 28 / 32 warps per SM
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python tools/c4_ab_lib.py full cur t128 p80 p80s p72 p64 cur > gpurun_out/c4_ab_park.log 2>&1; cat gpurun_out/c4_ab_park.log
