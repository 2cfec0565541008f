#!/bin/bash
# Round-2 GPU job 26: C4 A/B — leaf-lane ballots removed from k_wf_step_pt (noll) against the committed build (cur: no L2 window on the 32 MB nodes)
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python tools/c4_ab_lib.py full cur noll cur noll > gpurun_out/c4_ab_noll.log 2>&1; cat gpurun_out/c4_ab_noll.log
