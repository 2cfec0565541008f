"""Kernel-granularity probe on the GPU box: C4 (1M spheres) / C2 / C3 at reduced frames, warp chunks vs persistent lanes.
    python tools/pt_probe.py [lib.so]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from raytracing_renderer_cuda_b200 import capi
if len(sys.argv) > 1:
    capi.LIB_PATH = Path(sys.argv[1])
import raytracing_renderer_cuda_b200 as rt
ctx = rt.Context(0)
cases = [("c4", "random_spheres", dict(n=1_000_000), (1920, 1080, 4)), ("c2", "book1_final", {}, (1920, 1080, 32)),
         ("c3", "perlin_motion", {}, (1200, 600, 64))]
only = os.environ.get("PT_CASES", "c4,c2,c3").split(",")
for key, name, kw, (w, h, spp) in cases:
    if key not in only:
        continue
    sc = rt.Scene(ctx, rt.SceneDesc.builtin(name, **kw))
    ref = None
    sweep = [("warp", 0, 0)] + [("pt", r, l) for r in (16, 24) for l in (1, 4, 8, 12)]
    if os.environ.get("PT_SWEEP") == "short":
        sweep = [("warp", 0, 0), ("pt", 16, 1), ("pt", 16, 8)]
    for grain, refill, leaf in sweep:
        os.environ["RT_WF_GRAIN"] = grain
        os.environ["RT_PT_REFILL"] = str(max(refill, 1))
        os.environ["RT_PT_LEAF_LANES"] = str(max(leaf, 1))
        best = 1e9
        for _ in range(3):
            img, st = sc.render(rt.default_params(width=w, height=h, spp=spp))
            best = min(best, st.ms_total)
        if ref is None:
            ref = img.copy()
        import numpy as np
        diff = float(np.abs(img - ref).max())
        print(key, grain, refill, leaf, "ms", round(best, 2), "Mrays/s", round(st.rays / best / 1e3, 1), "iters", st.iterations,
              "maxdiff_vs_warp", diff, flush=True)
    sc.close()
