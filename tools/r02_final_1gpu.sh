#!/bin/bash
# Round-2 single-GPU measurement job (run through gpurun): GPU suite, smoke, bench lines of every config with the reference arms,
# launch list of the bench command, ncu --set full captures of the step kernels.  Outputs land in gpurun_out/r02f_*.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
RT_PARITY_LOG=$O/r02f_parity.jsonl timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > $O/r02f_pytest_gpu.log 2>&1; tail -n 3 $O/r02f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02f_smoke.log 2>&1; tail -n 1 $O/r02f_smoke.log
timeout 900 python bench.py > $O/r02f_bench_c1.json 2> $O/r02f_bench_c1.err; cut -c1-300 $O/r02f_bench_c1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02f_bench_ref_c1.json 2> $O/r02f_bench_ref_c1.err; cut -c1-300 $O/r02f_bench_ref_c1.json
for c in c2 c3 c4 c5; do timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline > $O/r02f_bench_$c.json 2> $O/r02f_bench_$c.err; cut -c1-200 $O/r02f_bench_$c.json; done
for c in c2 c3 c4; do :; done # (the reference arms of C2-C4 were measured by the previous run of this job: profiles/r02_bench_ref_c*.json)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02f_launches_c1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $O/r02f_ncu_list.log 2>&1; tail -n 1 $O/r02f_ncu_list.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_step_cta -s 4 -c 1 -o $O/r02f_prof_c1_cta -f python tools/c1_once.py > $O/r02f_ncu_c1.log 2>&1; tail -n 1 $O/r02f_ncu_c1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_tail -c 4 -o $O/r02f_prof_c1_tail -f python tools/c1_once.py > $O/r02f_ncu_c1t.log 2>&1; tail -n 1 $O/r02f_ncu_c1t.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_step_warp -s 3 -c 1 -o $O/r02f_prof_c2_warp -f python tools/c2_once.py > $O/r02f_ncu_c2.log 2>&1; tail -n 1 $O/r02f_ncu_c2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_step_warp -s 3 -c 1 -o $O/r02f_prof_c3_warp -f python tools/c2_once.py perlin_motion > $O/r02f_ncu_c3.log 2>&1; tail -n 1 $O/r02f_ncu_c3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_step_pt -s 2 -c 1 -o $O/r02f_prof_c4_pt -f python tools/c4_small.py > $O/r02f_ncu_c4.log 2>&1; tail -n 1 $O/r02f_ncu_c4.log
ls -la $O | grep r02f
