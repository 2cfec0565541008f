"""C4 (1M spheres) at a small frame: wavefront kernel granularities against the megakernel (same Philox paths)."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
ctx = rt.Context(0)
sc = rt.Scene(ctx, rt.SceneDesc.builtin("random_spheres", n=int(os.environ.get("C4_N", 1_000_000))))
w, h, spp = [int(v) for v in os.environ.get("C4_FRAME", "960,540,2").split(",")]
ref, st = sc.render_accum(rt.default_params(width=w, height=h, spp=spp, pipeline=capi.RT_PIPE_MEGAKERNEL))
print("mega rays", st.rays, "ms", round(st.ms_total, 1), flush=True)
for grain, refill, extra in [("warp", 1, {}), ("pt", 8, {}), ("pt", 24, {}), ("pt", 8, {"RT_NO_L2_PERSIST": "1"}), ("warp", 1, {"RT_NO_L2_PERSIST": "1"})]:
    os.environ["RT_WF_GRAIN"] = grain
    os.environ["RT_PT_REFILL"] = str(refill)
    os.environ.pop("RT_NO_L2_PERSIST", None)
    os.environ.update(extra)
    for rep in range(2):
        img, st = sc.render_accum(rt.default_params(width=w, height=h, spp=spp))
    d = np.abs(img - ref).max(axis=2)
    bad = np.argwhere(d > 1e-4)
    print(grain, refill, extra, "rays", st.rays, "ms", round(st.ms_total, 2), "Mrays/s", round(st.rays / st.ms_total / 1e3, 1),
          "pixels differing", len(bad), "max", float(d.max()), "count-channel diff", float(np.abs(img[..., 3] - ref[..., 3]).max()), flush=True)
    for y, x in bad[:5]:
        print("   ", int(y), int(x), img[y, x], ref[y, x])
