#!/bin/bash
# Round-2 GPU job 10 (gpurun --gpus 8): the rt_group path at 4 and 8 ranks (bit-exact sums), rt_multi with 8 devices,
# bench at N = 8 and 4 (weak C1 headline + c5 + strong sub-records).
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 > gpurun_out/pytest_multi8.log 2>&1; tail -n 12 gpurun_out/pytest_multi8.log
timeout 300 ./apps/render_scene --scene book1_final --width 1920 --height 1080 --spp 256 --gpus 8 --out gpurun_out/multi8.jpg > gpurun_out/app_multi8.log 2>&1; cat gpurun_out/app_multi8.log
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err; tail -n 3 gpurun_out/r02_bench_n$n.err | cut -c1-300; cut -c1-300 gpurun_out/r02_bench_n$n.json
done
