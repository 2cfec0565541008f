"""Device JPEG writer timing: encodes a synthetic frame at several sizes, prints device ms and GB/s of pixels."""
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
for (w, h) in [(1200, 600), (1920, 1080), (3840, 2160), (7680, 4320)]:
    img = oa.jpeg_test_image("photo", w, h, seed=1)
    dev = torch.from_numpy(img).cuda()
    out = np.empty(ctx.lib.rt_jpeg_max_bytes(w, h), np.uint8)
    best = 1e9
    for _ in range(5):
        n, ms = capi.jpeg_encode_device(ctx, dev.data_ptr(), w, h, 100, out)
        best = min(best, ms)
    print(f"{w}x{h}: {best:.3f} ms on the device, file {n} bytes, {w*h*3/best/1e6:.1f} GB/s of rgb8, {n/best/1e6:.2f} GB/s of stream", flush=True)

# ---- reader: host Huffman half (wall clock) and device pixel half (CUDA events)
import io, time
from PIL import Image
cases = []
earth = ROOT / "oracle" / "_ref" / "textures" / "earth.jpg"
if earth.exists():
    cases.append(("earth.jpg (reference texture, progressive 4:2:0)", earth.read_bytes()))
for (w, h) in [(1920, 1080), (3840, 2160)]:
    img = oa.jpeg_test_image("photo", w, h, seed=2)
    b = io.BytesIO()
    Image.fromarray(img).save(b, "JPEG", quality=90, subsampling=2)
    cases.append((f"{w}x{h} baseline 4:2:0 q90", b.getvalue()))
for name, data in cases:
    t0 = time.perf_counter()
    c = capi.jpeg_parse(data)
    t_host = (time.perf_counter() - t0) * 1e3
    capi.jpeg_coefficients_free(c)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        px, ms = capi.jpeg_decode(ctx, data)
        t_all = (time.perf_counter() - t0) * 1e3
        best = min(best, ms)
    print(f"decode {name}: file {len(data)} bytes -> {px.shape}; host Huffman half {t_host:.2f} ms, device pixel half {best:.3f} ms "
          f"(incl. H2D of the coefficients), whole call {t_all:.2f} ms", flush=True)
