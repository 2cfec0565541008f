"""Where do the frames of a random scene (tests/test_gpu_random_scenes.py) differ from the oracle's?  One shading step, ray by ray,
grouped by the material / texture of the sphere that was hit."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
from tests.oracle_api import Oracle, camera_rays, secondary_rays
from tests.test_gpu_random_scenes import random_document

seed, n = int(sys.argv[1]), int(sys.argv[2])
d = rt.SceneDesc.from_json(random_document(seed, n))
orc = Oracle().scene(d)
ctx = rt.Context(0)
sc = rt.Scene(ctx, d)
rays = camera_rays(d, 200_000, seed=7)
first = orc.trace(rays, arith=1)
rays = np.concatenate([rays, secondary_rays(d, first, seed=8)[:200_000]])
p = rt.default_params(tmin=1e-3)
want = orc.shade_probe(rays, p, arith=1)
got = sc.shade_probe(rays, p, use_bvh=True)
print("ids equal", np.array_equal(got["id"], want["id"]), "t equal", np.array_equal(got["t"], want["t"]))
sph = d.spheres()
id2mat = {int(s["id"]): int(s["material"]) for s in sph}
hit = want["id"] != capi.RT_INVALID_ID
mat = np.array([id2mat.get(int(i), -1) if h else -1 for i, h in zip(want["id"], hit)])
for m in sorted(set(mat[mat >= 0])):
    sel = mat == m
    mm = d.desc.materials[m]
    tk = d.desc.textures[mm.texture].kind if mm.texture >= 0 else -1
    da = np.abs(got["attenuation"][sel] - want["attenuation"][sel]).max(axis=1)
    de = np.abs(got["emitted"][sel] - want["emitted"][sel]).max(axis=1)
    fl = (got["continues"][sel] != want["continues"][sel])
    both = (got["continues"][sel] == 1) & (want["continues"][sel] == 1)
    dd = np.abs(got["scattered"]["direction"][sel][both] - want["scattered"]["direction"][sel][both]).max(axis=1) if both.any() else np.zeros(1)
    do = np.abs(got["scattered"]["origin"][sel][both] - want["scattered"]["origin"][sel][both]).max(axis=1) if both.any() else np.zeros(1)
    print(f"mat {m:2d} kind {mm.kind} tex {mm.texture:2d} texkind {tk:2d} n {sel.sum():6d}  att>1e-3 {float((da > 1e-3).mean()):.4f} max {da.max():.2e}  "
          f"emit max {de.max():.2e}  flips {float(fl.mean()):.5f}  dir>1e-3 {float((dd > 1e-3).mean()):.5f} max {dd.max():.2e}  origin max {do.max():.2e}")
