"""k_wf_tail (the thin end of a frame finished by CTA-local wavefronts) against the per-iteration step kernels alone: the
same paths — equal ray and sample counts, equal sums up to the order of the float atomics — on every scene kind, on small
frames (which run almost entirely in the tail kernel), on frames that switch to it in mid-flight, with emitter sampling,
and with a path pool much smaller than the frame (many generations of slots before the tail)."""
import numpy as np
import pytest

import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
from tests.conftest import SCENES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return rt.Context(0)


def _render(sc, p, monkeypatch, tail):
    monkeypatch.setenv("RT_WF_TAIL_PATHS", str(tail))
    return sc.render_accum(p)


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("size", [(64, 36, 4), (320, 180, 16)])
@pytest.mark.parametrize("flags", [0, capi.RT_RENDER_EMITTER_SAMPLING])
def test_tail_kernel_traces_the_same_paths(ctx, scene_descs, monkeypatch, name, size, flags):
    w, h, spp = size
    sc = rt.Scene(ctx, scene_descs[name])
    p = rt.default_params(width=w, height=h, spp=spp, flags=flags)
    ref, st_ref = _render(sc, p, monkeypatch, 0)  # step kernels only
    assert np.array_equal(ref[..., 3], np.full((h, w), spp, np.float32))
    for tail in (1000, 65536, 1 << 22):  # a late switch, the default, (nearly) the whole frame in the tail kernel
        got, st = _render(sc, p, monkeypatch, tail)
        assert st.paths == st_ref.paths and st.rays == st_ref.rays, (tail, st.rays, st_ref.rays)
        assert np.array_equal(got[..., 3], ref[..., 3]), tail
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-5), tail
        assert st.iterations <= st_ref.iterations


def test_tail_kernel_shortens_the_reference_frame(ctx, scene_descs, monkeypatch):
    """C1 at the reference's size: the frame with the tail kernel needs fewer launches and traces the same rays."""
    sc = rt.Scene(ctx, scene_descs["earth_emitter"])
    p = rt.default_params(width=1200, height=600, spp=20)
    ref, st_ref = _render(sc, p, monkeypatch, 0)
    monkeypatch.delenv("RT_WF_TAIL_PATHS")
    got, st = sc.render_accum(p)  # the default threshold
    assert st.rays == st_ref.rays and st.paths == st_ref.paths
    assert np.array_equal(got[..., 3], ref[..., 3])
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-5)
    assert st.launches < st_ref.launches


def test_deep_paths_finish_in_the_tail_kernel(ctx, monkeypatch):
    """A closed mirror-like scene: every path runs to the depth limit, the last ones inside k_wf_tail."""
    d = rt.SceneDesc.builtin("book1_final")
    sc = rt.Scene(ctx, d)
    p = rt.default_params(width=96, height=54, spp=4, max_depth=100)
    ref, st_ref = _render(sc, p, monkeypatch, 0)
    got, st = _render(sc, p, monkeypatch, 1 << 20)
    assert st.rays == st_ref.rays
    assert np.array_equal(got[..., 3], ref[..., 3]) and np.allclose(got, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", SCENES)
def test_tail_kernel_after_many_slot_generations(scene_descs, monkeypatch, name):
    """A 4 Ki-slot pool under a 147 k-path frame: 36 generations of slots, then the tail (a context of its own: the pool of
    a context only grows)."""
    monkeypatch.setenv("RT_WF_POOL", "4096")
    small = rt.Context(0)
    sc = rt.Scene(small, scene_descs[name])
    p = rt.default_params(width=96, height=48, spp=32)
    ref, st_ref = _render(sc, p, monkeypatch, 0)
    got, st = _render(sc, p, monkeypatch, 131072)
    assert st.paths == st_ref.paths == 96 * 48 * 32 and st.rays == st_ref.rays
    assert np.array_equal(got[..., 3], ref[..., 3]) and np.allclose(got, ref, rtol=1e-5, atol=1e-5)
    assert st.iterations < st_ref.iterations
