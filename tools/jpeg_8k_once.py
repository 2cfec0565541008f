"""One 8K device JPEG encode (for ncu --set full of the writer's kernels)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
from tests import oracle_api as oa
ctx = rt.Context(0)
w, h = 7680, 4320
img = oa.jpeg_test_image("photo", w, h, seed=1)
dev = torch.from_numpy(img).cuda()
out = np.empty(ctx.lib.rt_jpeg_max_bytes(w, h), np.uint8)
for _ in range(2):
    n, ms = capi.jpeg_encode_device(ctx, dev.data_ptr(), w, h, 100, out)
print(n, ms)
