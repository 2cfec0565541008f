"""Generates tests/golden/ref_cpu_golden.npz from the REFERENCE's own headers compiled for the host
(oracle/_ref/libref_cpu.so, built from /root/reference by `make -C oracle ref`).  Run in the build
container (needs /root/reference); the .npz it writes is what travels.

    python tests/golden/make_cpu_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import raytracing_renderer_cuda_b200 as rt  # noqa: E402
from raytracing_renderer_cuda_b200.assets import load_earth  # noqa: E402
from tests.oracle_api import RefCpu, camera_rays, secondary_rays  # noqa: E402

R = RefCpu()
rng = np.random.default_rng(20261018)
out = {}

# Perlin: SURVEY Appendix A points, lattice points, random points in a wide range
pts = [(0, 0, 0), (.5, .5, .5), (1.25, -2.5, 3.75), (-.3, .7, 10.1), (123.456, 7.89, -.12), (.1, .2, .3), (1, 2, 3),
       (-4, 255, 256), (-0.0001, 0.9999, 511.5)]
pts = np.array(pts + list(rng.uniform(-300, 300, size=(200, 3))) + list(rng.uniform(-2, 2, size=(200, 3))), np.float32)
out["perlin_p"] = pts
out["perlin_noise"] = np.array([R.perlin_noise(p) for p in pts], np.float32)
out["perlin_turb"] = np.array([R.turbulence(p) for p in pts], np.float32)

# optics
v = rng.normal(size=(256, 3)).astype(np.float32) * rng.uniform(0.1, 5, size=(256, 1)).astype(np.float32)
n = rng.normal(size=(256, 3)).astype(np.float32)
n /= np.linalg.norm(n, axis=1, keepdims=True)
mu = np.where(rng.random(256) < 0.5, np.float32(1.5), np.float32(1 / 1.5)).astype(np.float32)
out["opt_v"], out["opt_n"], out["opt_mu"] = v, n, mu
out["opt_reflect"] = np.array([R.reflect(a, b) for a, b in zip(v, n)], np.float32)
rr = [R.refract(a, b, float(m)) for a, b, m in zip(v, n, mu)]
out["opt_refract_ok"] = np.array([r[0] for r in rr], np.uint8)
out["opt_refract"] = np.array([r[1] for r in rr], np.float32)
cs = rng.random(256, dtype=np.float32)
out["opt_cos"] = cs
out["opt_shlick"] = np.array([R.shlick(float(c), 1.5) for c in cs], np.float32)

earth = load_earth()
for name, kw in (("earth_emitter", dict(image=earth)), ("book1_final", {}), ("perlin_motion", {})):
    d = rt.SceneDesc.builtin(name, **kw)
    rs = R.scene(d, use_bvh=False)
    rays = camera_rays(d, 2048, seed=11)
    hits = rs.trace(rays)
    sec = secondary_rays(d, hits, seed=12)[:2048]
    out[f"{name}_rays"] = np.concatenate([rays, sec])
    out[f"{name}_hits"] = np.concatenate([hits, rs.trace(sec)])
    mean, fb, nrays = rs.render(48, 24, 4, seed=1000, nthreads=4)
    out[f"{name}_fb_48x24x4"] = fb
    out[f"{name}_nrays_48x24x4"] = np.array([nrays], np.uint64)
    # textures: every texture of the scene at surface points of the first hits
    nt = d.desc.n_textures
    ok = hits["id"] != 0xFFFFFFFF
    P = hits["p"][ok][:24]
    U, V = hits["u"][ok][:24], hits["v"][ok][:24]
    tv = np.zeros((min(nt, 64), len(P), 3), np.float32)
    for t in range(tv.shape[0]):
        for k in range(len(P)):
            tv[t, k] = rs.texture_value(t, float(U[k]), float(V[k]), P[k])
    out[f"{name}_tex_p"], out[f"{name}_tex_u"], out[f"{name}_tex_v"], out[f"{name}_tex_value"] = P, U, V, tv
    cr = np.concatenate([rs.camera_ray(float(a), float(b), int(s)) for a, b, s in
                         zip(rng.random(16), rng.random(16), rng.integers(1, 1 << 30, 16))])
    out[f"{name}_camray_in"] = None  # placeholder replaced below

# camera rays need their inputs stored too: regenerate deterministically
rng2 = np.random.default_rng(7)
for name, kw in (("earth_emitter", dict(image=earth)), ("book1_final", {}), ("perlin_motion", {})):
    d = rt.SceneDesc.builtin(name, **kw)
    rs = R.scene(d, use_bvh=False)
    st = np.stack([rng2.random(16), rng2.random(16)], 1).astype(np.float32)
    seeds = rng2.integers(1, 1 << 30, 16).astype(np.uint64)
    out[f"{name}_camray_in"] = st
    out[f"{name}_camray_seed"] = seeds
    out[f"{name}_camray_out"] = np.concatenate([rs.camera_ray(float(a), float(b), int(s)) for (a, b), s in zip(st, seeds)])

path = ROOT / "tests" / "golden" / "ref_cpu_golden.npz"
np.savez_compressed(path, **out)
print("wrote", path, path.stat().st_size, "bytes")
