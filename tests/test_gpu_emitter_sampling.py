"""RT_RENDER_EMITTER_SAMPLING on the device, through the C-ABI: every kernel that implements the flag
(megakernel, the three wavefront granularities) against the oracle's restatement with the same random numbers, the
flagged estimator against the reference estimator (same expectation, less noise), and the flag's no-op cases."""
import numpy as np
import pytest

import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
from tests.test_emitter_sampling import DARK, lit_room_desc

pytestmark = pytest.mark.gpu
NEE = capi.RT_RENDER_EMITTER_SAMPLING


@pytest.fixture(scope="module")
def ctx():
    return rt.Context(0)


@pytest.mark.parametrize("pipe", [capi.RT_PIPE_WAVEFRONT, capi.RT_PIPE_MEGAKERNEL])
@pytest.mark.parametrize("bvh", ["none", "sah"])
def test_flagged_render_matches_oracle_same_random_numbers(ctx, oracle, pipe, bvh):
    d = lit_room_desc(bvh)
    w, h, spp = 96, 48, 2
    p = rt.default_params(width=w, height=h, spp=spp, pipeline=pipe, flags=NEE, max_depth=6, **DARK)
    got, st = rt.Scene(ctx, d).render_accum(p)
    want, nrays = oracle.scene(d).render(p, sampler=1, arith=1)
    assert st.paths == w * h * spp and np.array_equal(got[..., 3], np.full((h, w), spp, np.float32))
    assert abs(int(st.rays) - int(nrays)) <= nrays // 100
    # Same Philox blocks, same formulas: the paths coincide, so most pixels agree to rounding.  The rest are the r = 1000
    # floor's self-intersections (tmin = 1e-5 < ulp(1000), SURVEY.md 8a' item 2) flipping with the last ulp of a direction
    # — about 5 % of the PATHS of this scene with and without the flag (round-1 debugging run).
    diff = np.abs(got[..., :3] - want[..., :3]).max(axis=2)
    assert (diff < 1e-5).mean() > 0.75, float((diff < 1e-5).mean())
    mg, mw = float(got[..., :3].mean()), float(want[..., :3].mean())
    assert abs(mg - mw) / mw < 0.02, (mg, mw)
    # and the shadow rays are there: the flag traces more rays and lights the floor
    plain, st0 = rt.Scene(ctx, d).render_accum(rt.default_params(width=w, height=h, spp=spp, pipeline=pipe, max_depth=6, **DARK))
    assert st.rays > 1.2 * st0.rays


def test_every_wavefront_kernel_renders_the_flagged_frame(ctx, monkeypatch):
    """CTA chunks, warp chunks, persistent lanes and the megakernel trace the same shadow rays (Philox block (bounce, 2) of
    the same pixel/sample key) and drop the same emitter hits: identical ray counts, identical sums up to float order."""
    d = lit_room_desc("sah")
    sc = rt.Scene(ctx, d)
    w, h, spp = 160, 80, 8
    ref, st_ref = sc.render_accum(rt.default_params(width=w, height=h, spp=spp, pipeline=capi.RT_PIPE_MEGAKERNEL, flags=NEE, **DARK))
    for grain in ("cta", "warp", "pt"):
        monkeypatch.setenv("RT_WF_GRAIN", grain)
        got, st = sc.render_accum(rt.default_params(width=w, height=h, spp=spp, flags=NEE, **DARK))
        assert st.rays == st_ref.rays, grain
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-5), grain


def test_same_expectation_less_noise_on_the_device(ctx):
    sc = rt.Scene(ctx, lit_room_desc())
    w, h, spp = 128, 64, 512

    def render(flags, seed):
        acc, _ = sc.render_accum(rt.default_params(width=w, height=h, spp=spp, seed=seed, flags=flags, max_depth=8, **DARK))
        return acc[..., :3].astype(np.float64) / spp

    plain = [render(0, s) for s in (1, 2)]
    guided = [render(NEE, s) for s in (3, 4)]
    lit = np.maximum(guided[0], guided[1]).max(axis=2) < 0.6  # not the pixels that see a lamp (tests/test_emitter_sampling.py)
    assert lit.mean() > 0.9
    mse_plain = np.mean((plain[0] - plain[1])[lit] ** 2)
    mse_guided = np.mean((guided[0] - guided[1])[lit] ** 2)
    assert mse_guided < 0.2 * mse_plain, (mse_guided, mse_plain)
    mp, mg = np.mean((plain[0] + plain[1])[lit]) / 2, np.mean((guided[0] + guided[1])[lit]) / 2
    assert abs(mp - mg) / mp < 0.01, (mp, mg)
    diff = np.mean(((plain[0] + plain[1]) - (guided[0] + guided[1]))[lit] ** 2 / 4)
    assert diff < 1.5 * (mse_plain + mse_guided) / 4 + 1e-7, (diff, mse_plain, mse_guided)


def test_reference_scene_converges_to_the_same_frame(ctx, earth):
    """Config C1's scene (two emitters, one of them textured -> Q_EMIT): flag on and off meet at high sample counts.
    (Its sky is as bright as its emitters, so the flag buys nothing there — 1.4 x the variance in the oracle — but the
    expectation is the same: two independent 2048-spp frames of this size are ~49 dB apart.)"""
    sc = rt.Scene(ctx, rt.SceneDesc.builtin("earth_emitter", earth))
    w, h, spp = 200, 100, 2048
    a, _ = sc.render(rt.default_params(width=w, height=h, spp=spp))
    b, _ = sc.render(rt.default_params(width=w, height=h, spp=spp, flags=NEE))
    assert rt.psnr(a, b) > 44.0, rt.psnr(a, b)


def test_flag_without_emitters_and_default_flag(ctx):
    d = rt.SceneDesc.builtin("book1_final")
    sc = rt.Scene(ctx, d)
    a, sa = sc.render_accum(rt.default_params(width=96, height=54, spp=4))
    b, sb = sc.render_accum(rt.default_params(width=96, height=54, spp=4, flags=NEE))  # no emitter: the plain kernels run
    assert sa.rays == sb.rays and np.allclose(a, b, rtol=1e-6, atol=1e-6)
    with pytest.raises(capi.RtError):
        sc.render_accum(rt.default_params(width=8, height=8, spp=1, flags=0x80))
