"""The reference's own committed output of its hard-coded scene — renders/earth_emitter.jpg (README.md:17-18; 1200x600,
100 spp, JPEG quality 100), one of the two fixtures SURVEY.md 8c finds for this path — against the oracle (CPU) and the
CUDA pipelines (GPU).  The fixture is its 4x4 box filter, tests/golden/ref_render_earth_emitter_300x150.png
(tests/golden/make_ref_render_fixture.py).  A noise-limited gate, not a bit-exact one: the reference image was
made with cuRAND's XORWOW sequences on other hardware; what has to agree is the scene, the camera, the estimator,
the pixel finalisation and the writer's flip/quantisation.  Measured with the oracle: 31.6 dB at 300x150x16 spp,
36.9 dB at 300x150x64 spp, 44.4 dB at 1200x600x25 spp and 47.7 dB at the reference's own 1200x600x100 spp after the
same box filter."""
import numpy as np
import pytest
from PIL import Image

import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
from tests.conftest import ROOT

FIXTURE = ROOT / "tests" / "golden" / "ref_render_earth_emitter_300x150.png"


def _fixture() -> np.ndarray:
    return np.asarray(Image.open(FIXTURE).convert("RGB"), dtype=np.float64) / 255.0


def _written_bytes(rgb_bottom_up: np.ndarray) -> np.ndarray:
    """What the reference's writer makes of a finalised frame (main.cu:476-487): Y flip, int(255.999f * c) & 255."""
    q = (np.float32(255.999) * rgb_bottom_up[::-1].astype(np.float32)).astype(np.int32) & 255
    return q.astype(np.float64) / 255.0


def _box(img: np.ndarray, f: int) -> np.ndarray:
    h, w, _ = img.shape
    return img.reshape(h // f, f, w // f, f, 3).mean(axis=(1, 3))


def _psnr(a, b) -> float:
    return float(10.0 * np.log10(1.0 / np.mean((a - b) ** 2)))


@pytest.mark.parametrize("sampler", [0, 1])  # 0: the oracle's XORWOW-per-pixel stand-in, 1: the product's Philox keys
def test_oracle_reproduces_the_reference_render(oracle, earth, sampler):
    fix = _fixture()
    desc = rt.SceneDesc.builtin("earth_emitter", earth)
    spp = 64
    acc, _ = oracle.scene(desc).render(rt.default_params(width=300, height=150, spp=spp), sampler=sampler, arith=0, nthreads=8)
    img = np.sqrt(np.clip(acc[..., :3] / spp, 0.0, 1.0))  # main.cu:124-127
    got = _psnr(_written_bytes(img), fix)
    assert got >= 35.0, got  # measured 36.9 dB: the noise of 64 spp against 100 spp x 16 pixels
    # a wrong convention is far below the gate: not flipping the rows costs more than 15 dB
    assert _psnr(_written_bytes(img[::-1]), fix) < got - 15.0


@pytest.mark.gpu
@pytest.mark.parametrize("pipeline", [capi.RT_PIPE_WAVEFRONT, capi.RT_PIPE_MEGAKERNEL])
def test_cuda_frame_reproduces_the_reference_render(earth, pipeline):
    """The reference's frame itself — 1200x600, 100 spp, depth 50 (config C1) — through rt_render, then the writer
    conversion and the fixture's box filter."""
    fix = _fixture()
    ctx = rt.Context(0)  # (as the module-scoped fixtures of the other GPU tests: released with the object)
    scene = rt.Scene(ctx, rt.SceneDesc.builtin("earth_emitter", earth))
    img, st = scene.render(rt.default_params(width=1200, height=600, spp=100, pipeline=pipeline))
    assert st.paths == 1200 * 600 * 100
    got = _psnr(_box(_written_bytes(img), 4), fix)
    assert got >= 42.0, got  # the oracle (product sampler, same frame) reaches 47.7 dB; 44.4 dB with a quarter of the samples
    # and the library's own writer conversion is the one restated above
    assert np.array_equal(capi.quantize_rgb8(img).astype(np.float64) / 255.0, _written_bytes(img))
    scene.close()
