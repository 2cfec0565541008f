"""JPEG reader, CPU side: the product's HOST half (marker parsing + baseline/progressive Huffman decoding,
rt_jpeg_parse — plain host code, no device needed) followed by the oracle's pixel stages (oracle/jpeg_decode_oracle.cpp)
must reproduce what the reference's vendored stb_image.h decodes — committed fixtures, and the live library where
oracle/_ref has been built from /root/reference (there also the reference's own progressive textures/earth.jpg)."""
import io

import numpy as np
import pytest

from raytracing_renderer_cuda_b200 import capi
from tests import oracle_api as oa
from tests.conftest import ROOT

EARTH_JPG = ROOT / "oracle" / "_ref" / "textures" / "earth.jpg"


def _decode_cpu(data: bytes) -> np.ndarray:
    c = capi.jpeg_parse(data)
    try:
        return oa.oracle_jpeg_pixels(c)
    finally:
        capi.jpeg_coefficients_free(c)


@pytest.fixture(scope="module")
def golden():
    return np.load(ROOT / "tests" / "golden" / "jpeg_decode_golden.npz")


def test_host_half_plus_oracle_pixels_equal_stb_fixtures(golden):
    n = 0
    for key in sorted(k for k in golden.files if k.startswith("file")):
        idx = key[4:]
        got = _decode_cpu(golden[key].tobytes())
        want = golden[f"pix{idx}"]
        assert got.shape == want.shape and np.array_equal(got, want), str(golden[f"name{idx}"])
        n += 1
    assert n >= 50


def test_coefficient_planes_describe_the_file(golden):
    for key in sorted(k for k in golden.files if k.startswith("file")):
        name = str(golden[f"name{key[4:]}"])
        c = capi.jpeg_parse(golden[key].tobytes())
        cc = c.contents
        try:
            assert cc.progressive == int("progressive=True" in name)
            assert cc.n_comp == (1 if "grey" in name else 3)
            if "subsampling=2" in name or "4:2:0" in name:
                assert (cc.h_max, cc.v_max) == (2, 2) and (cc.comp[1].h, cc.comp[1].v) == (1, 1)
            for k in range(cc.n_comp):
                p = cc.comp[k]
                assert p.w2 % 8 == 0 and p.h2 % 8 == 0 and p.blocks_w * 8 == p.w2 and p.x <= p.w2 and p.y <= p.h2
        finally:
            capi.jpeg_coefficients_free(c)


@pytest.mark.parametrize("bad", [b"", b"\xff\xd8", b"not a jpeg at all", b"\xff\xd8\xff\xc0\x00\x05\x08", b"\xff\xd8\xff\xdb\x00\x03\x00\xff\xd9"])
def test_malformed_files_are_errors(bad):
    with pytest.raises(capi.RtError) as e:
        capi.jpeg_parse(bad)
    assert e.value.status == capi.RT_ERR_INVALID_ARG


def test_truncated_file_does_not_crash(golden):
    data = golden["file010"].tobytes()
    for cut in (len(data) // 3, len(data) // 2, len(data) - 3):
        try:
            c = capi.jpeg_parse(data[:cut])
            capi.jpeg_coefficients_free(c)  # like stb, a scan that ends early decodes what is there
        except capi.RtError:
            pass


def test_frame_size_that_overflows_an_int_is_rejected_like_stb(golden):
    """stb_image.h:3223 rejects width * height * components > INT_MAX ("too large"); the same check keeps a damaged
    frame header (65535 x 65535 on a 1 x 1 file) from sizing 25 GB of coefficient arrays — that used to end in a
    std::bad_alloc thrown through the C ABI."""
    import struct

    data = bytearray(golden["file000"].tobytes())
    sof = max(data.find(b"\xff\xc0"), data.find(b"\xff\xc2"))
    assert sof > 0
    data[sof + 5:sof + 9] = struct.pack(">HH", 65535, 65535)
    with pytest.raises(capi.RtError) as e:
        capi.jpeg_parse(bytes(data))
    assert e.value.status == capi.RT_ERR_INVALID_ARG and "too large" in str(e.value)
    if oa.REFSTB_SO.exists():
        with pytest.raises(Exception):
            oa.ref_stb_load_jpeg(bytes(data))


@pytest.mark.skipif(not oa.REFSTB_SO.exists(), reason="oracle/_ref/libref_stb.so is built from /root/reference only")
def test_pinned_live_against_stb_in_every_mode():
    from PIL import Image

    k = 0
    for (w, h) in [(9, 1), (1, 9), (16, 16), (31, 47), (129, 65), (250, 130)]:
        img = oa.jpeg_test_image("photo" if k % 2 else "noise", w, h, seed=300 + k)
        k += 1
        for prog in (False, True):
            for sub in (0, 1, 2):
                for q in (97, 40):
                    b = io.BytesIO()
                    Image.fromarray(img).save(b, "JPEG", quality=q, progressive=prog, subsampling=sub, optimize=bool(k % 2))
                    assert np.array_equal(_decode_cpu(b.getvalue()), oa.ref_stb_load_jpeg(b.getvalue())), (w, h, prog, sub, q)
        assert np.array_equal(_decode_cpu(oa.oracle_jpeg(img, 100)), oa.ref_stb_load_jpeg(oa.oracle_jpeg(img, 100)))


@pytest.mark.skipif(not (EARTH_JPG.exists() and oa.REFSTB_SO.exists()), reason="needs the reference's earth.jpg (oracle/_ref)")
def test_reference_texture_decodes_to_the_bytes_stb_decodes():
    """textures/earth.jpg (progressive, 4:2:0, 1200x600): stb == this reader == the committed assets/earth_stb.png."""
    from PIL import Image

    data = EARTH_JPG.read_bytes()
    got = _decode_cpu(data)
    assert np.array_equal(got, oa.ref_stb_load_jpeg(data))
    assert np.array_equal(got, np.asarray(Image.open(ROOT / "assets" / "earth_stb.png").convert("RGB")))


def _strip_segments(data: bytes, marker: int) -> bytes:
    """removes every segment with the given marker from the header part of a JPEG file (up to the first SOS)"""
    out, i = bytearray(data[:2]), 2
    while i + 4 <= len(data) and data[i] == 0xFF:
        m, ln = data[i + 1], (data[i + 2] << 8) | data[i + 3]
        if m == 0xDA:
            break
        if m != marker:
            out += data[i:i + 2 + ln]
        i += 2 + ln
    return bytes(out + data[i:])


def test_scan_that_selects_an_undefined_huffman_table_is_rejected(golden):
    """the tables were uninitialised until a DHT segment filled them, so a file whose scan names a table that
    was never defined decoded against stack garbage (out-of-bounds reads in the code-length search).  Such a file is now
    an error, for baseline and progressive files alike, and the outcome is deterministic."""
    tried = 0
    for key in list(golden.files)[:12]:
        if not key.startswith("file"):
            continue
        data = golden[key].tobytes()
        cut = _strip_segments(data, 0xC4)
        assert len(cut) < len(data)
        for _ in range(2):
            with pytest.raises(capi.RtError) as e:
                capi.jpeg_parse(cut)
            assert e.value.status == capi.RT_ERR_INVALID_ARG and "huffman" in str(e.value)
        tried += 1
    assert tried >= 4
