"""Image-texture ingest and the reference's second scene on the GPU through the C-ABI: the
environment-sphere scene of populate_scene_hdr (main.cu:136-182) against the oracle, with an image whose size is
neither the render size nor a multiple of anything convenient."""
import numpy as np
import pytest

import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
from tests.oracle_api import camera_rays, secondary_rays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return rt.Context(0)


def test_closest_hits_bit_exact(ctx, oracle, scene_descs):
    d = scene_descs["hdr_sphere"]
    rays = camera_rays(d, 100_000, seed=41)
    want = oracle.scene(d).trace(rays, arith=1)
    sec = secondary_rays(d, want, seed=42)  # start on the balls or on the inside of the r = 10 environment sphere
    want2 = oracle.scene(d).trace(sec, arith=1)
    sc = rt.Scene(ctx, d)
    assert sc.info().n_nodes == 0  # hitable_list without a bvh_node: the brute-force loop (hitable_list.h:66-78)
    assert (want["id"] != capi.RT_INVALID_ID).all()  # the camera sits inside the environment sphere: nothing escapes
    for r, w in ((rays, want), (sec, want2)):
        got = sc.trace_primary(r, use_bvh=False)
        assert np.array_equal(got["id"], w["id"]) and np.array_equal(got["t"], w["t"])
        assert np.array_equal(got["p"], w["p"]) and np.array_equal(got["n"], w["n"])


@pytest.mark.parametrize("pipe", [capi.RT_PIPE_WAVEFRONT, capi.RT_PIPE_MEGAKERNEL])
def test_first_bounce_image_lookup_matches_oracle(ctx, oracle, scene_descs, pipe):
    """max_depth = 1: camera ray, closest hit, and for the 94 % of the frame that sees the environment sphere the
    image look-up through get_sphere_uv (sphere.h:61-83, texture.h:118-132) with a 250x130 image."""
    d = scene_descs["hdr_sphere"]
    w, h, spp = 120, 60, 8
    p = rt.default_params(width=w, height=h, spp=spp, pipeline=pipe, max_depth=1)
    got, st = rt.Scene(ctx, d).render_accum(p)
    want, nrays = oracle.scene(d).render(p, sampler=1, arith=1)
    assert int(st.rays) == int(nrays) == w * h * spp
    diff = np.abs(got[..., :3] - want[..., :3]).max(axis=2)
    # a texel boundary can flip with the last ulp of atan2f/asinf (device libm vs host libm): a handful of samples
    assert np.median(diff) == 0.0 and (diff > 1e-4).mean() < 0.02, (float(np.median(diff)), float((diff > 1e-4).mean()))


@pytest.mark.parametrize("pipe", [capi.RT_PIPE_WAVEFRONT, capi.RT_PIPE_MEGAKERNEL])
def test_render_matches_oracle_same_random_numbers(ctx, oracle, scene_descs, pipe):
    d = scene_descs["hdr_sphere"]
    w, h, spp = 96, 48, 16
    p = rt.default_params(width=w, height=h, spp=spp, pipeline=pipe)
    got, st = rt.Scene(ctx, d).render_accum(p)
    want, nrays = oracle.scene(d).render(p, sampler=1, arith=1)
    assert st.paths == w * h * spp and np.array_equal(got[..., 3], np.full((h, w), spp, np.float32))
    assert abs(int(st.rays) - int(nrays)) <= nrays // 100
    diff = np.abs(got[..., :3] - want[..., :3]).max(axis=2) / spp
    assert np.median(diff) < 1e-2
    mg, mw = float(got[..., :3].mean()), float(want[..., :3].mean())
    assert abs(mg - mw) / mw < 0.01, (mg, mw)
    assert rt.psnr(oracle.tonemap(got), oracle.tonemap(want)) > 20.0


def test_grey_image_through_the_channel_conversion(ctx, oracle):
    """A one-channel image (stbi_loadf with ch == 1) ingested through rt_image_to_rgb renders like its RGB replica."""
    from tests.conftest import make_env_image

    env = make_env_image(64, 40)
    grey = env.mean(axis=2, keepdims=True).astype(np.float32)
    a = rt.SceneDesc.builtin("hdr_sphere", capi.image_to_rgb(grey))
    b = rt.SceneDesc.builtin("hdr_sphere", np.repeat(grey, 3, axis=2))
    p = rt.default_params(width=64, height=32, spp=4)
    ia, _ = rt.Scene(ctx, a).render_accum(p)
    ib, _ = rt.Scene(ctx, b).render_accum(p)
    assert np.array_equal(ia[..., 3], ib[..., 3])
    assert np.allclose(ia, ib, rtol=2e-6, atol=0)  # same paths; only the order of the float additions per pixel may differ
