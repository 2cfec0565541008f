#!/bin/bash
# Round-2 GPU job 37 (gpurun --gpus 8): config C5 in full through the C++ host (rt_multi_*): book-1 final scene, 7680x4320, 4096 spp,
# samples sharded over 8 B200, fused NVLink reduce + finalisation, JPEG of the 8K frame; a second run right after it (warm pools).
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for k in 1 2; do
timeout 600 ./apps/render_scene --scene book1_final --width 7680 --height 4320 --spp 4096 --gpus 8 --out gpurun_out/c5_full.jpg > gpurun_out/c5_full_$k.log 2>&1; cat gpurun_out/c5_full_$k.log
done
ls -la gpurun_out/c5_full.jpg
python - <<'PY'
from PIL import Image
im = Image.open("gpurun_out/c5_full.jpg"); print(im.size); im.resize((960, 540), Image.LANCZOS).save("gpurun_out/c5_full_960x540.jpg", quality=92)
PY
rm -f gpurun_out/c5_full.jpg
