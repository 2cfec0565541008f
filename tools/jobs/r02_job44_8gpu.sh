#!/bin/bash
# Round-2 GPU job 44 (gpurun --gpus 8): the multi-GPU tests at 2 / 4 / 8 members and the N = 8 bench line with the one-pixel-per-thread reduce
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 500 > gpurun_out/r02f_pytest_multi8.log 2>&1; tail -n 3 gpurun_out/r02f_pytest_multi8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02f_bench_n8.json 2> gpurun_out/r02f_bench_n8.err; tail -n 2 gpurun_out/r02f_bench_n8.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/r02f_bench_n8.json')); c=d['c5']; print('N=8 c1', round(d['value']), round(d['ms_per_step'],3), 'reduce', round(d['ms_reduce_finalise'],3), '| c5', round(c['value']), round(c['ms_per_step'],2), 'reduce', round(c['ms_reduce_finalise'],3), 'GB/s', round(c['nvlink_gbs_per_gpu'],1), '| strong', {k:(round(v['value']), round(v['ms_per_step'],3), round(v['ms_reduce_finalise'],3)) for k,v in d['strong'].items()})"
timeout 300 ./apps/render_scene --scene book1_final --width 7680 --height 4320 --spp 64 --gpus 8 --out gpurun_out/m8.ppm 2>&1 | tail -n 1; rm -f gpurun_out/m8.ppm
