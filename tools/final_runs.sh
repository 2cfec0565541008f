#!/bin/bash
# Round-end measurement job for one GPU (run through gpurun): tests, smoke, benches of every config, reference arms,
# launch list of the bench command and ncu captures.  Outputs land in gpurun_out/.
set -x
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
for c in c2 c3 c4 c5; do python bench.py --config $c --steps 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_c1.json 2> gpurun_out/bench_ref_c1.err
for c in c2 c3 c4; do python bench.py --impl reference --config $c --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_ref_$c.json 2> gpurun_out/bench_ref_$c.err; done
python tools/jpeg_probe.py > gpurun_out/jpeg_probe.log 2>&1
python tools/nee_probe.py > gpurun_out/nee_probe.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_wf_step -s 4 -c 1 -o gpurun_out/prof_wf_c1_final -f python tools/c1_once.py > gpurun_out/ncu_c1_final.log 2>&1
cut -c1-300 gpurun_out/bench_c*.json gpurun_out/bench_ref_c*.json
