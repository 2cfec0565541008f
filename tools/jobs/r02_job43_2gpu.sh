#!/bin/bash
# Round-2 GPU job 43 (gpurun --gpus 2): pixels per thread of the fused reduce with peers (RT_REDUCE_G=4 / 1) at 8K, two members:
# rt_multi through the C++ app (event-timed reduce + finalisation) and the c5 sub-record of the N = 2 bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for g in 4 1 4 1; do
RT_REDUCE_G=$g timeout 300 ./apps/render_scene --scene book1_final --width 7680 --height 4320 --spp 8 --gpus 2 --out gpurun_out/m2.ppm 2>&1 | tail -n 1 | sed "s/^/G=$g /"
done
rm -f gpurun_out/m2.ppm
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 250 2>&1 | tail -n 2
for g in 4 1; do
RT_REDUCE_G=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2962$g bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_g$g.json 2> gpurun_out/bench_n2_g$g.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_g$g.json')); c=d['c5']; print('G=$g N=2 c1 ms', round(d['ms_per_step'],3), 'reduce', round(d['ms_reduce_finalise'],3), '| c5 ms', round(c['ms_per_step'],2), 'reduce', round(c['ms_reduce_finalise'],3), 'GB/s', round(c['nvlink_gbs_per_gpu'],1), '| strong c5', round(d['strong']['c5']['ms_per_step'],2), round(d['strong']['c5']['ms_reduce_finalise'],3))"
done
