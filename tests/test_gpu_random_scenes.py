"""Differential parity on RANDOM scene documents (seeded): the three fixed benchmark scenes cannot exercise every
combination of material, texture, motion and acceleration structure, so this file draws scenes — 3 to 300 spheres,
every material and procedural texture kind of the reference, moving spheres, a ground, defocus — and requires of each:
  * closest hits through the list, the host SAH tree and the GPU LBVH tree bit-identical to the oracle (ids, t),
  * the megakernel, the wavefront step kernels alone (RT_WF_TAIL_PATHS=0), the wavefront with the tail kernel, and every
    acceleration structure trace the same paths (equal ray counts, equal sums up to float-add order),
  * the frame equal to the oracle's with the same random numbers at tmin = 1e-3 (no acne flips), pixel for pixel in the bulk."""
import json

import numpy as np
import pytest

import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
from tests.conftest import record_parity
from tests.oracle_api import camera_rays, secondary_rays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return rt.Context(0)


def random_document(seed: int, n: int) -> str:
    r = np.random.default_rng(seed)
    f = lambda lo, hi: float(r.uniform(lo, hi))
    col = lambda: [f(0.05, 0.95), f(0.05, 0.95), f(0.05, 0.95)]
    textures = {
        "c0": {"type": "constant", "color": col()}, "c1": {"type": "constant", "color": col()},
        "perlin": {"type": "noise", "noise": "PERLIN", "density": f(1, 8)},
        "turb": {"type": "noise", "noise": "TURBULANCE", "density": f(1, 6)},
        "marble": {"type": "noise", "noise": "MARBLE", "density": f(0.5, 4)},
        "wood": {"type": "wood", "color1": col(), "color2": col(), "density": f(0.5, 4), "hardness": f(5, 60)},
        "glow": {"type": "constant", "color": [f(0.5, 1), f(0.5, 1), f(0.5, 1)]},
    }
    textures["check"] = {"type": "checker", "even": "c0", "odd": "marble"}
    textures["check2"] = {"type": "checker", "even": "check", "odd": "wood"}  # a checker of a checker (texture.h:41-48)
    materials = {}
    for k, t in enumerate(["c0", "c1", "perlin", "turb", "marble", "wood", "check", "check2"]):
        materials[f"lam{k}"] = {"type": "lambertian", "texture": t}
    for k in range(3):
        materials[f"met{k}"] = {"type": "metal", "albedo": col(), "roughness": [0.0, f(0.05, 0.5), f(0.5, 1.0)][k]}
    materials["glass"] = {"type": "dielectric", "ri": f(1.3, 1.8), "tint": [1, 1, 1]}
    materials["tinted"] = {"type": "dielectric", "ri": 1.5, "tint": col()}
    materials["lamp"] = {"type": "emitter", "texture": "glow", "intensity": f(1, 4)}
    materials["lamp2"] = {"type": "emitter", "texture": "perlin", "intensity": f(1, 3)}  # a textured emitter (queue Q_EMIT)
    names = list(materials)
    extent = 1.5 + 0.35 * n ** (1 / 3)
    objects = [{"type": "sphere", "center": [0, -500.4, 0], "radius": 500, "material": str(r.choice(["lam4", "lam6", "lam0"]))}]
    for k in range(n - 1):
        c = [f(-extent, extent), f(-0.3, 0.6 * extent), f(-extent, extent)]
        rad = f(0.08, 0.45)
        m = str(r.choice(names))
        if r.random() < 0.15:
            d = [f(-0.4, 0.4), f(-0.2, 0.4), f(-0.4, 0.4)]
            objects.append({"type": "moving_sphere", "center0": c, "center1": [c[i] + d[i] for i in range(3)], "time0": 0, "time1": 1,
                            "radius": rad, "material": m})
        else:
            objects.append({"type": "sphere", "center": c, "radius": rad, "material": m})
    cam = {"lookfrom": [f(-1, 1) * extent, f(0.5, 1.5) * extent, 3.0 * extent], "lookat": [0, 0.2, 0], "up": [0, 1, 0], "vfov": f(25, 45),
           "aspect": 4 / 3, "aperture": f(0, 0.15), "focus_dist": "auto", "time0": 0, "time1": float(r.choice([0.0, 1.0]))}
    return json.dumps({"camera": cam, "textures": textures, "materials": materials, "objects": objects, "bvh": "auto"})


CASES = [(101, 3), (102, 9), (103, 12), (104, 13), (105, 40), (106, 300)]


@pytest.mark.parametrize("seed,n", CASES)
def test_random_scene_closest_hits_match_the_oracle(ctx, oracle, seed, n):
    d = rt.SceneDesc.from_json(random_document(seed, n))
    orc = oracle.scene(d)
    rays = camera_rays(d, 60_000, seed=seed)
    want = orc.trace(rays, arith=1)
    sec = secondary_rays(d, want, seed=seed + 1)
    want2 = orc.trace(sec, arith=1)
    assert (want["id"] != capi.RT_INVALID_ID).mean() > 0.3  # the camera sees the scene
    for mode in (capi.RT_BVH_NONE, capi.RT_BVH_HOST_SAH, capi.RT_BVH_GPU_LBVH):
        d.set_bvh_mode(mode)
        sc = rt.Scene(ctx, d)
        for rr, w in ((rays, want), (sec, want2)):
            got = sc.trace_primary(rr, use_bvh=mode != capi.RT_BVH_NONE)
            assert np.array_equal(got["id"], w["id"]), (seed, mode)
            assert np.array_equal(got["t"], w["t"]), (seed, mode)
    d.set_bvh_mode(capi.RT_BVH_AUTO)


@pytest.mark.parametrize("seed,n", CASES)
def test_random_scene_every_pipeline_traces_the_same_paths(ctx, seed, n, monkeypatch):
    d = rt.SceneDesc.from_json(random_document(seed, n))
    w, h, spp = 96, 72, 6
    ref = st_ref = None
    for mode in (capi.RT_BVH_NONE, capi.RT_BVH_HOST_SAH, capi.RT_BVH_GPU_LBVH):
        d.set_bvh_mode(mode)
        sc = rt.Scene(ctx, d)
        for pipe, tail in ((capi.RT_PIPE_MEGAKERNEL, None), (capi.RT_PIPE_WAVEFRONT, 0), (capi.RT_PIPE_WAVEFRONT, 1000), (capi.RT_PIPE_WAVEFRONT, None)):
            if tail is None:
                monkeypatch.delenv("RT_WF_TAIL_PATHS", raising=False)
            else:
                monkeypatch.setenv("RT_WF_TAIL_PATHS", str(tail))
            got, st = sc.render_accum(rt.default_params(width=w, height=h, spp=spp, pipeline=pipe, tmin=1e-3))
            assert np.array_equal(got[..., 3], np.full((h, w), spp, np.float32)), (seed, mode, pipe, tail)
            if ref is None:
                ref, st_ref = got, st
            else:
                assert st.rays == st_ref.rays, (seed, mode, pipe, tail, st.rays, st_ref.rays)
                assert np.allclose(got, ref, rtol=1e-5, atol=1e-5), (seed, mode, pipe, tail)
    d.set_bvh_mode(capi.RT_BVH_AUTO)


@pytest.mark.parametrize("seed,n", CASES)
@pytest.mark.parametrize("max_depth", [3, 50])
def test_random_scene_frame_matches_the_oracle_with_the_same_random_numbers(ctx, oracle, seed, n, max_depth):
    """Measured on a B200 (profiles/r02_parity.md): at full depth 0.03-0.5 % of the pixels differ for up to 40 spheres and 11.8 %
    in the 300-sphere scene, whose quarter of mirrors and glass balls amplifies the last bit of an SFU sincos by the curvature
    of every reflection (ray counts still agree to 0.2 %, frame means to 1e-3); cut after three bounces it is 1.9 %.  The
    yardstick is tests/test_random_scene_sensitivity.py: the ORACLE against itself with round-to-nearest instead of
    round-toward-zero vector operators differs in 8.7 % / 0.55 % of the same scene's pixels and in <= 0.2 % of the others' —
    the scene is that sensitive to one ulp, whoever computes it; one shading step agrees ray by ray for every material of
    it (tools/random_scene_probe.py: attenuation within 2.4e-5, no discrete flips, hits bit-identical)."""
    d = rt.SceneDesc.from_json(random_document(seed, n))
    w, h, spp = 64, 48, 6
    p = rt.default_params(width=w, height=h, spp=spp, tmin=1e-3, max_depth=max_depth)
    got, st = rt.Scene(ctx, d).render_accum(p)
    want, nrays = oracle.scene(d).render(p, sampler=1, arith=1)
    diff = np.abs(got[..., :3] - want[..., :3]).max(axis=2) / spp
    frac = float((diff > 1e-3).mean())
    psnr = rt.psnr(oracle.tonemap(got), oracle.tonemap(want))
    mean_rel = float(abs(got[..., :3].mean() - want[..., :3].mean()) / want[..., :3].mean())
    record_parity("random_scene_same_rng", seed=seed, n=n, max_depth=max_depth, size=f"{w}x{h}x{spp}", frac_gt_1e3=frac,
                  median=float(np.median(diff)), psnr_db=psnr, rays_gpu=int(st.rays), rays_oracle=int(nrays), mean_rel_diff=mean_rel)
    assert np.array_equal(got[..., 3], np.full((h, w), spp, np.float32))
    assert np.median(diff) < 1e-5
    assert frac < (0.02 if max_depth == 3 or n <= 40 else 0.25), frac
    assert abs(int(st.rays) - int(nrays)) <= max(8, nrays // 200)
    assert mean_rel < 5e-3, mean_rel  # the pixels that differ do so without a bias: the frame means agree
