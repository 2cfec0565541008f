"""The experimental barrier-free kernel (RT_WF_GRAIN=ring) and the hybrid tail (RT_WF_TAIL) against the default
wavefront kernels: every kernel traces the same paths (Philox keys), so sample counts and ray counts must be EQUAL and
colour sums equal up to the order of the float atomics.  OPT-IN — `RT_TEST_EXPERIMENTAL=1 python -m pytest
tests/test_gpu_experimental_ring.py -m gpu`: the emitter-sampling instantiations, the edge-case frames and the whole
RT_WF_TAIL path had not run on a GPU when this was written (profiles/r01_ring.md), and neither mode is selected by
default, so the round's GPU suite does not depend on them."""
import os

import numpy as np
import pytest

import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
from tests.test_emitter_sampling import DARK, lit_room_desc

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.environ.get("RT_TEST_EXPERIMENTAL"), reason="opt-in: RT_TEST_EXPERIMENTAL=1"),
              pytest.mark.timeout(120)]

MODES = [("RT_WF_GRAIN", "ring"), ("RT_WF_TAIL", "4096"), ("RT_WF_TAIL", "1000000")]


@pytest.fixture(scope="module")
def ctx():
    return rt.Context(0)


def _same_frame(sc, params, monkeypatch, mode):
    ref, st_ref = sc.render_accum(params)
    monkeypatch.setenv(*mode)
    for _ in range(2):  # the rings and their counters live across frames
        got, st = sc.render_accum(params)
        assert st.rays == st_ref.rays and st.paths == st_ref.paths, mode
        assert np.array_equal(got[..., 3], ref[..., 3]), mode
        assert np.allclose(got[..., :3], ref[..., :3], rtol=1e-5, atol=1e-4), mode
    monkeypatch.delenv(mode[0])
    return st


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name,size", [("earth_emitter", (96, 48, 8)), ("earth_emitter", (600, 300, 32)), ("perlin_motion", (320, 160, 16)),
                                       ("book1_final", (320, 180, 16)), ("hdr_sphere", (200, 100, 16))])
def test_same_frame_as_the_default_kernels(ctx, scene_descs, monkeypatch, mode, name, size):
    w, h, spp = size
    st = _same_frame(rt.Scene(ctx, scene_descs[name]), rt.default_params(width=w, height=h, spp=spp), monkeypatch, mode)
    if mode[0] == "RT_WF_GRAIN":
        assert st.launches == 3 and st.iterations == 1  # k_ring_fill, k_ring_commit, k_wf_ring


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("bvh", ["none", "sah"])
def test_emitter_sampling_instantiations(ctx, monkeypatch, mode, bvh):
    p = rt.default_params(width=160, height=80, spp=8, flags=capi.RT_RENDER_EMITTER_SAMPLING, **DARK)
    _same_frame(rt.Scene(ctx, lit_room_desc(bvh)), p, monkeypatch, mode)


@pytest.mark.parametrize("mode", MODES)
def test_edge_case_frames(ctx, scene_descs, monkeypatch, mode):
    sc = rt.Scene(ctx, scene_descs["earth_emitter"])
    for kw in (dict(width=33, height=7, spp=1), dict(width=1, height=1, spp=1), dict(width=64, height=32, spp=0),
               dict(width=64, height=32, spp=4, max_depth=0), dict(width=64, height=32, spp=4, max_depth=1),
               dict(width=64, height=32, spp=3, sample_offset=5)):
        _same_frame(sc, rt.default_params(**kw), monkeypatch, mode)


@pytest.mark.parametrize("mode", MODES)
def test_large_scene_through_the_lbvh(ctx, monkeypatch, mode):
    sc = rt.Scene(ctx, rt.SceneDesc.builtin("random_spheres", n=20000))
    _same_frame(sc, rt.default_params(width=320, height=180, spp=4), monkeypatch, mode)
