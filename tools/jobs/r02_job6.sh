#!/bin/bash
# Round-2 GPU job 6: profiles of k_wf_blk (why is it slower?) and of k_wf_step_cta in the j4 and cur libraries (same source
# of the shading code, 7.4 against 8.1 ms on one box), plus the A/B again with the in-tree library.
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
cp raytracing_renderer_cuda_b200/librt_b200.so gpurun_variants/librt_intree.so
AB_NO_MEGA=1 AB_CASES=c1 timeout 600 python tools/ab_test.py j4 cur intree j4 cur intree > gpurun_out/ab_place2.log 2>&1; cat gpurun_out/ab_place2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_blk -s 4 -c 1 -o gpurun_out/r02_prof_c1_blk -f python tools/c1_blk_once.py > gpurun_out/ncu_c1_blk.log 2>&1; tail -n 2 gpurun_out/ncu_c1_blk.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
for v in j4 cur; do timeout 300 ncu --metrics $M --clock-control none -k regex:k_wf_step -s 4 -c 3 --csv --log-file gpurun_out/ncu_cta_$v.csv python tools/c1_once.py gpurun_variants/librt_$v.so > gpurun_out/ncu_cta_$v.log 2>&1; done
tail -n 30 gpurun_out/ncu_cta_j4.csv | cut -c1-400
