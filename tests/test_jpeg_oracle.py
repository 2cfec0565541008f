"""Output stage, CPU side: the oracle of the JPEG writer (oracle/jpeg_oracle.cpp) against bytes written
by the reference's own vendored stb_image_write.h — the committed fixtures (tests/golden/jpeg_golden.npz, generated
by tests/golden/make_jpeg_golden.py) and, where oracle/_ref has been built from /root/reference, the live library.
Bar: byte-exact files."""
import io

import numpy as np
import pytest

from tests import oracle_api as oa
from tests.conftest import ROOT


@pytest.fixture(scope="module")
def golden():
    return np.load(ROOT / "tests" / "golden" / "jpeg_golden.npz")


def _cases(golden):
    for key in sorted(k for k in golden.files if k.startswith("case")):
        kind, w, h, q, seed = golden[key]
        yield key[4:], str(kind), int(w), int(h), int(q), int(seed)


def test_oracle_writes_the_bytes_stb_wrote(golden):
    n = 0
    for idx, kind, w, h, q, seed in _cases(golden):
        img = oa.jpeg_test_image(kind, w, h, seed)
        assert oa.oracle_jpeg(img, q) == golden[f"jpg{idx}"].tobytes(), (kind, w, h, q)
        n += 1
    assert n >= 16


def test_fixture_files_decode_to_their_input(golden):
    """The fixtures are real JPEG files: an independent decoder (Pillow/libjpeg) reads them back to the input."""
    from PIL import Image

    for idx, kind, w, h, q, seed in _cases(golden):
        if q < 100:
            continue
        img = oa.jpeg_test_image(kind, w, h, seed)
        dec = np.asarray(Image.open(io.BytesIO(golden[f"jpg{idx}"].tobytes())).convert("RGB"))
        assert dec.shape == img.shape
        assert np.abs(dec.astype(int) - img.astype(int)).max() <= 6, (kind, w, h)  # quantiser 1 + YCbCr rounding


@pytest.mark.skipif(not oa.REFSTB_SO.exists(), reason="oracle/_ref/libref_stb.so is built from /root/reference only")
def test_oracle_pinned_live_against_stb():
    k = 0
    for (w, h) in [(8, 8), (1, 1), (9, 1), (1, 9), (16, 16), (31, 47), (100, 50), (129, 65)]:
        for kind in ("noise", "smooth", "flat", "sat", "photo"):
            img = oa.jpeg_test_image(kind, w, h, seed=100 + k)
            for q in (100, 91, 90, 60, 25, 1, 0):
                assert oa.oracle_jpeg(img, q) == oa.ref_stb_jpeg(img, q), (kind, w, h, q)
            k += 1


def test_reference_frame_size_oracle_runs():
    """1200x600 at quality 100 (the reference's render.jpg, main.cu:491): structure checks on the oracle's file."""
    img = oa.jpeg_test_image("photo", 1200, 600, seed=7)
    f = oa.oracle_jpeg(img, 100)
    assert f[:2] == b"\xff\xd8" and f[-2:] == b"\xff\xd9"
    sof = f.index(b"\xff\xc0")
    assert f[sof + 5:sof + 9] == bytes([600 >> 8, 600 & 255, 1200 >> 8, 1200 & 255])
    assert f[sof + 11] == 0x11  # quality 100: no chroma subsampling (stb_image_write.h:1443-1444)
