"""RT_RENDER_EMITTER_SAMPLING (the reference README's roadmap item "Improve Sampling on emitter
objects", README.md:27-28), CPU side.  The reference has no implementation of it, so there is nothing to pin against:
the tests check the property that defines the feature — the flag changes the ESTIMATOR, not its expectation:

  * a lambertian hit's shadow ray, weighted by p_ref / p_sel, estimates exactly the part of the reference's scatter
    distribution d = n + ball (material.h:112) that points at an emitter sphere — direction AND length of d (the
    reference never normalises it, and tmin is in units of |d|);
  * oracle renders with and without the flag agree within Monte-Carlo noise, and the flagged estimator is the less
    noisy one where emitters light the scene;
  * the flag is off by default and does nothing in a scene without emitters.
"""
import json

import numpy as np
import pytest

import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi

# a lambertian floor and ball lit by two small emitters (one of them moving), a dark world colour
LIT_ROOM = {
    "camera": {"lookfrom": [0, 1.5, 6], "lookat": [0, 0.3, 0], "vfov": 30, "aspect": 2, "aperture": 0.0, "focus_dist": 6,
               "time0": 0, "time1": 1},
    "textures": {"grey": {"type": "constant", "color": [0.7, 0.7, 0.7]}, "blue": {"type": "constant", "color": [0.2, 0.3, 0.8]},
                 "warm": {"type": "constant", "color": [12, 9, 6]}, "cold": {"type": "constant", "color": [4, 8, 12]}},
    "materials": {"floor": {"type": "lambertian", "texture": "grey"}, "ball": {"type": "lambertian", "texture": "blue"},
                  "mirror": {"type": "metal", "albedo": [0.9, 0.9, 0.9], "roughness": 0.1},
                  "lamp": {"type": "emitter", "texture": "warm"}, "lamp2": {"type": "emitter", "texture": "cold", "intensity": 1.5}},
    "objects": [
        {"type": "sphere", "center": [0, -1000, 0], "radius": 1000, "material": "floor"},
        {"type": "sphere", "center": [0, 0.5, 0], "radius": 0.5, "material": "ball"},
        {"type": "sphere", "center": [1.3, 0.4, 0.3], "radius": 0.4, "material": "mirror"},
        {"type": "sphere", "center": [-1.2, 1.6, 0.5], "radius": 0.15, "material": "lamp"},
        {"type": "moving_sphere", "center0": [1.0, 1.8, -0.5], "center1": [1.6, 1.8, -0.5], "time0": 0, "time1": 1, "radius": 0.1,
         "material": "lamp2"},
    ],
    "bvh": "none",
}
DARK = dict(world=(0.02, 0.02, 0.03), bloom=0.0)


def lit_room_desc(bvh="none"):
    doc = dict(LIT_ROOM, bvh=bvh)
    return rt.SceneDesc.from_json(json.dumps(doc))


def _reference_samples(n_hat, count, rng):
    """d = n + uniform-in-unit-ball (material.h:112, utils.h:61-77) by rejection, in numpy."""
    pts = rng.uniform(-1, 1, size=(count * 2, 3))
    pts = pts[(pts * pts).sum(1) < 1.0][:count]
    return pts + n_hat


def test_flag_is_off_by_default_and_is_validated():
    p = rt.default_params()
    assert p.flags == 0
    assert capi.RT_RENDER_EMITTER_SAMPLING == 1
    capi.check_layout(rt.load_library())  # sizeof(rt_render_params) of the ctypes mirror == the library's


def _lamp_cones(p, time):
    """(axis, cos_max) of the two lamps of LIT_ROOM seen from p at `time` (the second lamp moves)."""
    out = []
    for c, r in ((np.array([-1.2, 1.6, 0.5]), 0.15), (np.array([1.0 + 0.6 * time, 1.8, -0.5]), 0.1)):
        v = c - np.asarray(p, np.float64)
        dist = np.linalg.norm(v)
        out.append((v / dist, -1.0 if dist <= r else np.sqrt(1 - (r / dist) ** 2)))
    return out


@pytest.mark.parametrize("p,n,time", [
    ((0.0, 0.0, 1.0), (0.0, 1.0, 0.0), 0.25),      # floor point, both lamps above the horizon
    ((0.0, 1.0, 0.0), (0.0, 1.0, 0.0), 0.9),       # top of the ball
    ((0.35, 0.5, 0.35), (0.70710678, 0.0, 0.70710678), 0.5),  # side of the ball: a lamp straddles the horizon
    ((-1.2, 1.55, 0.5), (0.0, 1.0, 0.0), 0.0),     # inside the warm lamp: its cone is the whole sphere
])
def test_shadow_ray_estimates_the_emitter_part_of_the_reference_scatter(oracle, p, n, time):
    sc = oracle.scene(lit_room_desc())
    N = 60000
    d = np.empty((N, 3), np.float32)
    w = np.empty(N, np.float64)
    for i in range(N):
        d[i], w[i] = sc.light_sample(p, n, time, seed=77, index=i)
    nh = np.asarray(n, np.float64)
    cones = _lamp_cones(p, time)
    live = w > 0
    length = np.linalg.norm(d[live].astype(np.float64), axis=1)
    unit = d[live].astype(np.float64) / length[:, None]
    cos_t = unit @ nh
    assert (cos_t > 0).all() and (length <= 2 * cos_t * (1 + 1e-5)).all()  # d ends inside the unit ball centred at n
    assert np.logical_or.reduce([unit @ ax >= cm - 1e-6 for ax, cm in cones]).all()  # every shadow ray aims at a lamp

    # reference: plain samples of n + ball, restricted to the directions inside a lamp's cone
    ref = _reference_samples(nh, 1_500_000, np.random.default_rng(3))
    rl = np.linalg.norm(ref, axis=1)
    ru = ref / rl[:, None]
    at_lamp = np.logical_or.reduce([ru @ ax >= cm for ax, cm in cones])
    dl = np.zeros(N)
    dc = np.zeros(N)
    dl[live], dc[live] = length, cos_t
    for f_ref, f_sel in ((np.ones(len(ref)), np.ones(N)), (rl, dl), ((ru @ nh) ** 2, dc ** 2)):
        want = at_lamp * f_ref
        est = w * f_sel
        tol = 5 * (est.std() / np.sqrt(N) + want.std() / np.sqrt(len(want))) + 1e-5
        assert abs(est.mean() - want.mean()) < tol, (est.mean(), want.mean(), tol)
    assert w.max() < 1.0 or cones[0][1] < 0  # weights are of the order of a lamp's solid angle, not of the path


def test_oracle_render_same_expectation_less_noise(oracle):
    sc = oracle.scene(lit_room_desc())
    w, h = 48, 24

    def render(flags, seed, spp):
        acc, _ = sc.render(rt.default_params(width=w, height=h, spp=spp, seed=seed, flags=flags, max_depth=8, **DARK), sampler=1, arith=1)
        assert np.array_equal(acc[..., 3], np.full((h, w), spp, np.float32))
        return acc[..., :3].astype(np.float64) / spp

    spp = 192
    plain = [render(0, s, spp) for s in (11, 12)]
    guided = [render(capi.RT_RENDER_EMITTER_SAMPLING, s, spp) for s in (21, 22)]
    # pixels that SEE a lamp (directly or in the mirror) are as noisy as the lamp's edge in both estimators; the flag is
    # about the rest of the frame — what the lamps light
    lit = np.maximum(guided[0], guided[1]).max(axis=2) < 0.6  # (the low-noise frames decide: no selection on outliers)
    assert lit.mean() > 0.9
    mse_plain = np.mean((plain[0] - plain[1])[lit] ** 2)    # = 2 x the estimator's variance at `spp` samples
    mse_guided = np.mean((guided[0] - guided[1])[lit] ** 2)
    assert mse_guided < 0.2 * mse_plain, (mse_guided, mse_plain)  # measured: 26 x less
    # same expectation: the difference of the two estimators is noise of the size the two variances predict
    mp, mg = np.mean((plain[0] + plain[1])[lit]) / 2, np.mean((guided[0] + guided[1])[lit]) / 2
    assert abs(mp - mg) / mp < 0.03, (mp, mg)
    diff = np.mean(((plain[0] + plain[1]) - (guided[0] + guided[1]))[lit] ** 2 / 4)
    assert diff < 1.5 * (mse_plain + mse_guided) / 4 + 1e-6, (diff, mse_plain, mse_guided)


def test_flag_does_nothing_without_emitters(oracle):
    d = rt.SceneDesc.builtin("book1_final")
    sc = oracle.scene(d)
    a, ra = sc.render(rt.default_params(width=24, height=12, spp=2), sampler=1, arith=1)
    b, rb = sc.render(rt.default_params(width=24, height=12, spp=2, flags=capi.RT_RENDER_EMITTER_SAMPLING), sampler=1, arith=1)
    assert ra == rb and np.array_equal(a, b)


def test_json_render_block_and_unknown_flags():
    p = rt.default_params()
    doc = dict(LIT_ROOM, render={"width": 64, "height": 32, "spp": 4, "emitter_sampling": 1})
    rt.SceneDesc.from_json(json.dumps(doc), params=p)
    assert (p.width, p.height, p.spp, p.flags) == (64, 32, 4, capi.RT_RENDER_EMITTER_SAMPLING)
    doc["render"]["emitter_sampling"] = 0
    rt.SceneDesc.from_json(json.dumps(doc), params=p)
    assert p.flags == 0
