"""C4 A/B on one box: 1 M spheres, 1920x1080x4 probe frame and the full 3840x2160x64 frame, node forms / env variants.
    python tools/c4_ab.py [full]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import raytracing_renderer_cuda_b200 as rt
ctx = rt.Context(0)
sc = rt.Scene(ctx, rt.SceneDesc.builtin("random_spheres", n=1_000_000))
print("scene", sc.info().n_spheres, "nodes", sc.info().n_nodes, flush=True)
frames = [(1920, 1080, 4)] + ([(3840, 2160, 64)] if len(sys.argv) > 1 else [])
ref = {}
for w, h, spp in frames:
    for label, env in (("float128", {"RT_BVH4": "f"}), ("quant64", {"RT_BVH4": "q"}), ("float128", {"RT_BVH4": "f"}), ("quant64", {"RT_BVH4": "q"})):
        os.environ.update(env)
        best = 1e9
        for _ in range(2 if spp > 4 else 3):
            img, st = sc.render(rt.default_params(width=w, height=h, spp=spp))
            best = min(best, st.ms_total)
        key = (w, h, spp)
        if key not in ref:
            ref[key] = img.copy()
        print(f"{w}x{h}x{spp} {label:9s} ms {best:9.2f}  Mrays/s {st.rays / best / 1e3:8.1f}  iters {st.iterations}  maxdiff {float(np.abs(img - ref[key]).max()):.2e}", flush=True)
