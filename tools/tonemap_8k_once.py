"""The framebuffer passes at 8K (7680x4320) for ncu: k_tonemap (float4 accumulator -> float RGB + flipped rgb8) and
k_reduce_tonemap with one member (the fused multi-GPU finalisation reading a local accumulator)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import raytracing_renderer_cuda_b200 as rt
W, H = 7680, 4320
ctx = rt.Context(0)
acc = torch.rand((H, W, 4), device="cuda") + 1.0
rgb = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
rgb8 = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    rt.tonemap_device(ctx, acc.data_ptr(), W, H, rgb.data_ptr(), rgb8.data_ptr())
    rt.reduce_tonemap_peers(ctx, [acc.data_ptr()], 0, W, H, 0, H, rgb.data_ptr(), rgb8.data_ptr())
ctx.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ev[0].record(); rt.tonemap_device(ctx, acc.data_ptr(), W, H, rgb.data_ptr(), rgb8.data_ptr())
ev[1].record(); rt.reduce_tonemap_peers(ctx, [acc.data_ptr()], 0, W, H, 0, H, rgb.data_ptr(), rgb8.data_ptr())
ev[2].record(); torch.cuda.synchronize()
b = W * H * (16 + 12 + 3)
print(f"k_tonemap {ev[0].elapsed_time(ev[1]):.3f} ms {b / ev[0].elapsed_time(ev[1]) / 1e6:.0f} GB/s | k_reduce_tonemap(1 member) {ev[1].elapsed_time(ev[2]):.3f} ms {b / ev[1].elapsed_time(ev[2]) / 1e6:.0f} GB/s")
