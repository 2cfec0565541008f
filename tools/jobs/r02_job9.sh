#!/bin/bash
# Round-2 GPU job 9: A/B — CTA kernel instantiated for the constant-bank list only (smaller code), block kernel at 80 registers
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
AB_NO_MEGA=1 AB_CASES=c1 timeout 600 python tools/ab_test.py div3inl listc blk80@RT_WF_GRAIN=blk div3inl listc blk80@RT_WF_GRAIN=blk > gpurun_out/ab_listc.log 2>&1; cat gpurun_out/ab_listc.log
