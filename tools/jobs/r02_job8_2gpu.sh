#!/bin/bash
# Round-2 GPU job 8 (gpurun --gpus 2): the multi-GPU entries of the C-ABI on real hardware.
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 > gpurun_out/pytest_multi2.log 2>&1; tail -n 25 gpurun_out/pytest_multi2.log
timeout 300 ./apps/render_scene --scene book1_final --width 640 --height 360 --spp 64 --gpus 2 --out gpurun_out/multi2.ppm > gpurun_out/app_multi2.log 2>&1; cat gpurun_out/app_multi2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -n 5 gpurun_out/r02_bench_n2.err | cut -c1-300; cut -c1-400 gpurun_out/r02_bench_n2.json
