#!/bin/bash
# After tools/r02_final_1gpu.sh has run through gpurun: copies the bench lines / logs of gpurun_out/r02f_* into profiles/ and
# regenerates the markdown summaries (launch list, parity values, ncu captures).  Runs without a GPU.
set -e
cd "$(dirname "$0")/.."
O=gpurun_out
for c in c1 c2 c3 c4 c5; do [ -s $O/r02f_bench_$c.json ] && cp $O/r02f_bench_$c.json profiles/r02_bench_$c.json; done
for c in c1 c2 c3 c4; do [ -s $O/r02f_bench_ref_$c.json ] && cp $O/r02f_bench_ref_$c.json profiles/r02_bench_ref_$c.json; done
cp $O/r02f_pytest_gpu.log profiles/r02_pytest_gpu.log
cp $O/r02f_launches_c1.csv profiles/r02_launches_c1.csv
python tools/launch_summary.py $O/r02f_launches_c1.csv "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs (round 2, final build)" r02_launches_c1.csv > profiles/r02_launches_c1.md
python tools/parity_report.py $O/r02f_parity.jsonl "r02 — measured values behind the parity gates (GPU suite of the final round-2 build on a B200, tools/r02_final_1gpu.sh)" > profiles/r02_parity.md
sum() { python tools/ncu_summary.py $O/$1.ncu-rep "$3" > profiles/$2.md; sed -i 's#`gpurun_out/\(r02f_[a-z0-9_]*\.ncu-rep\)`#`\1` (scratch file of the measurement job, not kept)#' profiles/$2.md; }
sum r02f_prof_c1_cta r02_c1_final_ncu "r02 — k_wf_step_cta<0,0,1> of the final round-2 build, C1, steady-state iteration (16 Mi live slots)"
sum r02f_prof_c1_tail r02_c1_tail_ncu "r02 — k_wf_tail<0,0,1> (CTA-local wavefronts, 512-thread CTAs), C1: the tail launches of one frame (the long one is the armed launch)"
sum r02f_prof_c2_warp r02_c2_final_ncu "r02 — k_wf_step_warp<1,0> of the final round-2 build (centre / half-extent nodes), C2 book-1 final (485 spheres, host SAH BVH), 1920x1080x32"
sum r02f_prof_c3_warp r02_c3_final_ncu "r02 — k_wf_step_warp<1,0> of the final round-2 build, C3 perlin_motion (145 primitives, host SAH BVH, Perlin / checker / wood textures, moving spheres), 1920x1080x32"
sum r02f_prof_c4_pt r02_c4_final_ncu "r02 — k_wf_step_pt<0,1> of the final round-2 build (64-byte nodes, sign-selected visit, 128 threads x 7 CTAs, no L2 window), 1 M spheres 1920x1080x4, iteration 2"
echo refreshed
