"""Output stage on the GPU, through the C-ABI: the CUDA JPEG writer must produce the file the
reference's stbi_write_jpg produces — byte-exact against the oracle (oracle/jpeg_oracle.cpp, itself pinned to the real
stb) and against the committed stb outputs; at frame sizes the oracle would take long for, through properties
(an independent decoder reads the file back; re-encoding is deterministic)."""
import io

import numpy as np
import pytest

import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
from tests import oracle_api as oa
from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return rt.Context(0)


def test_cuda_writes_the_bytes_stb_wrote(ctx):
    g = np.load(ROOT / "tests" / "golden" / "jpeg_golden.npz")
    for key in sorted(k for k in g.files if k.startswith("case")):
        kind, w, h, q, seed = g[key]
        img = oa.jpeg_test_image(str(kind), int(w), int(h), int(seed))
        got = capi.jpeg_encode(ctx, img, int(q))
        assert got == g[f"jpg{key[4:]}"].tobytes(), (kind, w, h, q)


@pytest.mark.parametrize("quality", [100, 95, 90, 50, 7])
def test_cuda_matches_oracle_on_ragged_and_edge_sizes(ctx, quality):
    k = 0
    for (w, h) in [(1, 1), (8, 8), (9, 7), (15, 17), (16, 16), (17, 33), (257, 3), (3, 257), (640, 360), (1201, 599)]:
        for kind in ("noise", "sat", "photo", "flat"):
            if w * h > 100000 and kind in ("sat", "flat"):
                continue
            img = oa.jpeg_test_image(kind, w, h, seed=500 + k)
            k += 1
            assert capi.jpeg_encode(ctx, img, quality) == oa.oracle_jpeg(img, quality), (kind, w, h, quality)


def test_reference_frame_byte_exact(ctx, scene_descs):
    """C1 through the whole device output stage: rt_render_jpeg == stb(quantise(flip(rt_render)))."""
    sc = rt.Scene(ctx, scene_descs["earth_emitter"])
    # one sample per pixel: every pixel is a single float addition, so two renders of the frame are bit-identical
    # (with more samples the order of the atomic adds may flip a last bit, and with it a byte of the file)
    p = rt.default_params(width=1200, height=600, spp=1)
    img, _ = sc.render(p)
    rgb8 = capi.quantize_rgb8(img)  # main.cu:475-488 on the host
    f, st = sc.render_jpeg(p, 100)
    assert f.tobytes() == oa.oracle_jpeg(rgb8, 100)
    assert st.paths == 1200 * 600 and st.ms_d2h > 0
    from PIL import Image

    dec = np.asarray(Image.open(io.BytesIO(f.tobytes())).convert("RGB"))
    assert dec.shape == (600, 1200, 3)
    assert capi.psnr(dec, rgb8, peak=255.0) > 45.0


def test_too_small_buffer_is_an_error_not_an_overrun(ctx):
    import ctypes as C

    img = oa.jpeg_test_image("noise", 64, 64, 1)
    out = np.full(4096 + 8, 0xAB, np.uint8)
    n = C.c_size_t(0)
    st = ctx.lib.rt_jpeg_encode(ctx._h, img.ctypes.data, 64, 64, 100, out.ctypes.data, 4096, C.byref(n))
    assert st == capi.RT_ERR_INVALID_ARG and n.value > 4096
    assert (out == 0xAB).all()
    with pytest.raises(capi.RtError):
        capi.jpeg_encode(ctx, np.zeros((0, 0, 3), np.uint8), 100)


def test_8k_frame_properties(ctx):
    """BASELINE config C5's frame (7680x4320): too slow for the scalar oracle in a unit test, so check what does not
    depend on it — the file decodes (libjpeg) to the input within the quantiser's error, a second run gives the same
    bytes, and a 16-row band equals the oracle's encoding of that band (MCU rows are independent up to DC prediction,
    so the band is encoded as its own image)."""
    import torch
    from PIL import Image

    Image.MAX_IMAGE_PIXELS = None
    w, h = 7680, 4320
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    rng = np.random.default_rng(3)
    img = np.stack([127 + 100 * np.sin(x / 97) * np.cos(y / 61), 127 + 90 * np.cos(x / 41 + y / 83), 60 + (x + y) / 100], -1)
    img = np.clip(img + rng.normal(0, 4, img.shape), 0, 255).astype(np.uint8)
    dev = torch.from_numpy(img).cuda()
    out = np.empty(ctx.lib.rt_jpeg_max_bytes(w, h), np.uint8)
    n1, ms = capi.jpeg_encode_device(ctx, dev.data_ptr(), w, h, 100, out)
    first = out[:n1].tobytes()
    n2, _ = capi.jpeg_encode_device(ctx, dev.data_ptr(), w, h, 100, out)
    assert n1 == n2 and out[:n2].tobytes() == first and ms > 0
    dec = np.asarray(Image.open(io.BytesIO(first)).convert("RGB"))
    assert dec.shape == img.shape and capi.psnr(dec, img, peak=255.0) > 45.0
    band = np.ascontiguousarray(img[:16])
    assert capi.jpeg_encode(ctx, band, 100) == oa.oracle_jpeg(band, 100)


def test_cpp_app_renders_a_json_scene_to_jpeg(ctx, scene_descs, tmp_path):
    """The C++ host program (apps/render_scene.cpp, the replacement of the reference's main()) with a JSON scene and the
    device output stage, against the same frame rendered through the Python binding."""
    import subprocess
    from PIL import Image

    app = ROOT / "apps" / "render_scene"
    out = tmp_path / "app.jpg"
    r = subprocess.run([str(app), "--scene", str(ROOT / "assets" / "scenes" / "earth_emitter.json"), "--width", "300", "--height", "150",
                        "--spp", "16", "--out", str(out)], capture_output=True, text=True, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "JPEG" in r.stdout
    got = np.asarray(Image.open(out).convert("RGB"))
    sc = rt.Scene(ctx, scene_descs["earth_emitter"])
    f, _ = sc.render_jpeg(rt.default_params(width=300, height=150, spp=16), 100)
    want = np.asarray(Image.open(io.BytesIO(f.tobytes())).convert("RGB"))
    assert got.shape == want.shape == (150, 300, 3)
    assert capi.psnr(got, want, peak=255.0) > 50.0  # same paths; float add order may flip a last bit before quantisation
