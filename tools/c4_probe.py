"""C4 (1M random spheres): render rate at a reduced frame for the BVH builders / kernel granularities."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = rt.SceneDesc.builtin("random_spheres", n=n)
ctx = rt.Context(0)
for mode, name in ((capi.RT_BVH_GPU_LBVH, "lbvh"), (capi.RT_BVH_HOST_SAH, "sah")):
    if name == "sah" and os.environ.get("NO_SAH"): continue
    d.set_bvh_mode(mode)
    sc = rt.Scene(ctx, d)
    i = sc.info()
    for k in range(2):
        img, st = sc.render(rt.default_params(width=1920, height=1080, spp=4))
    print(name, os.environ.get("RT_WF_GRAIN", "auto"), "build ms", round(i.ms_build, 2), "depth", i.bvh_depth, "render ms", round(st.ms_total, 1), "Mrays/s", round(st.rays / st.ms_total / 1e3, 1), "iters", st.iterations, flush=True)
    sc.close()
