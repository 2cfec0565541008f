#!/bin/bash
# Round-2 GPU job 46: the default bench command of HEAD, end to end (the pipelined e2e leg got an exception guard after the last bench run)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 600 python bench.py > gpurun_out/r02f_bench_c1.json 2> gpurun_out/r02f_bench_c1.err ) 2>&1 | grep real
tail -n 3 gpurun_out/r02f_bench_c1.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/r02f_bench_c1.json')); e=d['e2e']; print('C1', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(e['ms_per_step'],3), e['mode'], 'pipelined', e['pipelined'], 'serial', round(e['serial']['ms_per_step'],3), 'cpu', round(d['cpu_baseline']['value'],1), 'others', {k: round(v['value'],1) for k,v in d['other_configs'].items()}, 'launches', d['gpu_launches'], d['clocks'])"
