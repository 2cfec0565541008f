"""The oracle (oracle/rt_oracle.cpp, this repo's CPU restatement) against golden vectors recorded from
the reference's OWN headers compiled for the host (tests/golden/make_cpu_golden.py) and against the only
known-answer table the reference ships (sphere.h:71-77).  Bit-exact: both sides are IEEE float on x86
with true round-toward-zero vec3 arithmetic."""
import numpy as np
import pytest

from raytracing_renderer_cuda_b200 import capi
from tests.conftest import SCENES

import raytracing_renderer_cuda_b200 as rt


def test_uv_known_answers_from_reference_comment(oracle):
    # sphere.h:71-77
    table = {(1, 0, 0): (0.5, 0.5), (0, 1, 0): (0.5, 1.0), (0, 0, 1): (0.25, 0.5), (-1, 0, 0): (0.0, 0.5),
             (0, -1, 0): (0.5, 0.0), (0, 0, -1): (0.75, 0.5)}
    for n, (u, v) in table.items():
        gu, gv = oracle.sphere_uv(n)
        assert abs(gu - u) < 1e-6 and abs(gv - v) < 1e-6, (n, gu, gv)


def test_survey_appendix_a_perlin(oracle):
    # SURVEY.md Appendix A (reference headers, CPU shim)
    rows = [((0, 0, 0), 0.5, 0.0), ((.5, .5, .5), 0.375, 0.25), ((1.25, -2.5, 3.75), 0.5040183, 0.1330366),
            ((-.3, .7, 10.1), 0.4311016, 0.3028953), ((123.456, 7.89, -.12), 0.4876935, 0.1616516), ((.1, .2, .3), 0.6756146, 0.4188601)]
    for p, n, t in rows:
        assert abs(oracle.perlin_noise(p) - n) < 1e-6 and abs(oracle.turbulence(p) - t) < 1e-6


def test_perlin_bit_exact(oracle, cpu_golden):
    g = cpu_golden
    got_n = np.array([oracle.perlin_noise(p) for p in g["perlin_p"]], np.float32)
    got_t = np.array([oracle.turbulence(p) for p in g["perlin_p"]], np.float32)
    assert np.array_equal(got_n, g["perlin_noise"])
    assert np.array_equal(got_t, g["perlin_turb"])


def test_optics_bit_exact(oracle, cpu_golden):
    g = cpu_golden
    refl = np.array([oracle.reflect(a, b) for a, b in zip(g["opt_v"], g["opt_n"])], np.float32)
    assert np.array_equal(refl, g["opt_reflect"])
    rr = [oracle.refract(a, b, float(m)) for a, b, m in zip(g["opt_v"], g["opt_n"], g["opt_mu"])]
    assert np.array_equal(np.array([r[0] for r in rr], np.uint8), g["opt_refract_ok"])
    assert 0 < g["opt_refract_ok"].sum() < len(rr)  # both branches (refraction / total internal reflection) covered
    ok = g["opt_refract_ok"].astype(bool)
    assert np.array_equal(np.array([r[1] for r in rr], np.float32)[ok], g["opt_refract"][ok])
    sh = np.array([oracle.shlick(float(c), 1.5) for c in g["opt_cos"]], np.float32)
    assert np.array_equal(sh, g["opt_shlick"])


@pytest.mark.parametrize("name", SCENES)
def test_closest_hit_bit_exact(oracle, cpu_golden, scene_descs, name):
    g = cpu_golden
    rays = np.ascontiguousarray(g[f"{name}_rays"]).view(capi.RAY_DTYPE).reshape(-1)
    want = np.ascontiguousarray(g[f"{name}_hits"]).view(capi.HIT_DTYPE).reshape(-1)
    got = oracle.scene(scene_descs[name]).trace(rays, arith=0)
    for f in capi.HIT_DTYPE.names:  # id, t, p, n and u/v (incl. the stale u/v a moving sphere inherits)
        assert np.array_equal(got[f], want[f]), f
    assert (want["id"] != capi.RT_INVALID_ID).mean() > 0.5


@pytest.mark.parametrize("name", SCENES)
def test_textures_and_camera_bit_exact(oracle, cpu_golden, scene_descs, name):
    g = cpu_golden
    sc = oracle.scene(scene_descs[name])
    tv = g[f"{name}_tex_value"]
    for t in range(tv.shape[0]):
        for k, p in enumerate(g[f"{name}_tex_p"]):
            got = sc.texture_value(t, float(g[f"{name}_tex_u"][k]), float(g[f"{name}_tex_v"][k]), p)
            assert np.array_equal(got, tv[t, k]), (t, k)
    want = np.ascontiguousarray(g[f"{name}_camray_out"]).view(capi.RAY_DTYPE).reshape(-1)
    for k, ((a, b), s) in enumerate(zip(g[f"{name}_camray_in"], g[f"{name}_camray_seed"])):
        got = sc.camera_ray(float(a), float(b), int(s))
        assert got.tobytes() == want[k:k + 1].tobytes()


@pytest.mark.parametrize("name", SCENES)
def test_whole_render_bit_exact(oracle, cpu_golden, scene_descs, name):
    """color() + render with the reference's sampling: every draw, bounce and rounding must agree for
    the finished framebuffer and the ray count to be identical."""
    p = rt.default_params(width=48, height=24, spp=4)
    acc, nrays = oracle.scene(scene_descs[name]).render(p, sampler=0, arith=0, nthreads=4)
    assert nrays == int(cpu_golden[f"{name}_nrays_48x24x4"][0])
    assert np.array_equal(oracle.tonemap(acc), cpu_golden[f"{name}_fb_48x24x4"])
