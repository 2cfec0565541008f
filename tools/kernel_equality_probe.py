import os, sys, numpy as np
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
ctx = rt.Context(0)
sc = rt.Scene(ctx, rt.SceneDesc.builtin("random_spheres", n=1_000_000))
w, h, spp = 640, 360, 2
ref, sr = sc.render_accum(rt.default_params(width=w, height=h, spp=spp, pipeline=capi.RT_PIPE_MEGAKERNEL))
for grain in ("pt", "warp"):
    os.environ["RT_WF_GRAIN"] = grain
    for rep in range(2):
        got, st = sc.render_accum(rt.default_params(width=w, height=h, spp=spp))
        d = np.abs(got[..., :3] - ref[..., :3]).max(axis=2)
        print(os.environ.get("RT_B200_LIB", "default")[-16:], os.environ.get("RT_NO_L2_PERSIST", "-"), grain, "rays", st.rays, "ref", sr.rays, "iters", st.iterations,
              "count_equal", bool(np.array_equal(got[..., 3], ref[..., 3])), "count_sum", float(got[..., 3].sum()), "px_diff", int((d > 1e-5).sum()), flush=True)
