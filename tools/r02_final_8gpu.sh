#!/bin/bash
# Round-2 multi-GPU measurement job (gpurun --gpus 8): the multi-GPU tests on 8 GPUs, rt_multi through the C++ app, bench at N = 8, 4, 2
# (weak C1 headline + c5 + strong sub-records) with the final build.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 > gpurun_out/r02f_pytest_multi8.log 2>&1; tail -n 6 gpurun_out/r02f_pytest_multi8.log
timeout 300 ./apps/render_scene --scene book1_final --width 1920 --height 1080 --spp 256 --gpus 8 --out gpurun_out/r02f_multi8.jpg > gpurun_out/r02f_app_multi8.log 2>&1; cat gpurun_out/r02f_app_multi8.log
for n in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02f_bench_n$n.json 2> gpurun_out/r02f_bench_n$n.err; tail -n 3 gpurun_out/r02f_bench_n$n.err | cut -c1-300; cut -c1-300 gpurun_out/r02f_bench_n$n.json
done
