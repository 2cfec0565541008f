"""Prints how closely the CUDA render follows the oracle when both use the same random numbers."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200.assets import load_earth
from tests.oracle_api import Oracle
O = Oracle(); ctx = rt.Context(0)
for name, kw, (w, h, spp) in (("earth_emitter", dict(image=load_earth()), (96, 48, 16)), ("book1_final", {}, (64, 36, 8)), ("perlin_motion", {}, (80, 40, 8))):
    d = rt.SceneDesc.builtin(name, **kw)
    for depth in (1, 2, 3, 50):
        p = rt.default_params(width=w, height=h, spp=spp, max_depth=depth)
        got, st = rt.Scene(ctx, d).render_accum(p)
        want, nr = O.scene(d).render(p, sampler=1, arith=1)
        diff = np.abs(got[..., :3] - want[..., :3]).max(axis=2) / spp
        print(name, depth, "rays", st.rays, nr, "median", np.median(diff), "frac>1e-3", (diff > 1e-3).mean(), "frac>1e-5", (diff > 1e-5).mean(),
              "psnr", rt.psnr(O.tonemap(got), O.tonemap(want)), "mean rel", abs(got[..., :3].mean() - want[..., :3].mean()) / want[..., :3].mean(), flush=True)
