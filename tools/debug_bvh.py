import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
from tests.oracle_api import camera_rays, secondary_rays
ctx = rt.Context(0)
d = rt.SceneDesc.builtin("random_spheres", n=20_000)
sph = d.spheres()
def scene(mode):
    d.set_bvh_mode(mode); s = rt.Scene(ctx, d); d.set_bvh_mode(0); return s
rays = camera_rays(d, 100_000, seed=31)
L = scene(capi.RT_BVH_NONE)
lst = L.trace_primary(rays, use_bvh=False)
sec = secondary_rays(d, lst, seed=32)
lst2 = L.trace_primary(sec, use_bvh=False)
for mode in (2, 3):
    S = scene(mode)
    i = S.info(); print("mode", mode, "nodes", i.n_nodes, "depth", i.bvh_depth, "build ms", i.ms_build)
    for nm, r, w in (("cam", rays, lst), ("sec", sec, lst2)):
        g = S.trace_primary(r, use_bvh=True)
        bad = np.nonzero((g["id"] != w["id"]) | (g["t"] != w["t"]))[0]
        print(nm, "mismatches", len(bad))
        for b in bad[:6]:
            k = int(np.nonzero(sph["id"] == w["id"][b])[0][0]) if w["id"][b] != 0xFFFFFFFF else -1
            print("  ray", b, r[b], "\n   list", w["id"][b], w["t"][b], "bvh", g["id"][b], g["t"][b], "\n   sphere", sph[k] if k >= 0 else None)
