"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol include/rt_api.h
declares, structure layouts agree, the facade flattens the reference's scene graph as documented, the
writer conversion restates main.cu:475-488, and compute entries fail loudly without a device."""
import ctypes as C
import re

import numpy as np
import pytest

import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
from tests.conftest import ROOT


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "rt_api.h").read_text()
    declared = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", header))
    lib = capi.load_library()
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rt_api_version() == 1
    capi.check_layout(lib)


def test_no_cpu_fallback_without_device():
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(capi.RtError) as e:
        rt.Context(0)
    assert e.value.status == 5 and "no CPU fallback" in str(e.value)


def test_product_never_links_the_oracle():
    import subprocess

    out = subprocess.check_output(["ldd", str(capi.LIB_PATH)], text=True)
    assert "oracle" not in out and "ref_cpu" not in out
    for f in (ROOT / "raytracing_renderer_cuda_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".cpp", ".hpp"):
            txt = f.read_text()
            assert "liboracle" not in txt and "oracle_api" not in txt and "rt_oracle" not in txt, f


def test_builtin_scenes_through_the_facade(earth):
    c1 = rt.SceneDesc.builtin("earth_emitter", earth)
    d = c1.desc
    assert d.n_spheres == 8 and d.n_images == 1 and d.bvh_mode == capi.RT_BVH_AUTO
    s = c1.spheres()
    assert list(s["id"]) == list(range(8))                                 # set_id values, main.cu:200-315
    assert (s["flags"][7] & capi.RT_SPHERE_MOVING) and (s["flags"][2] & capi.RT_SPHERE_INSIDE)
    assert np.allclose(s["center0"][1], (0, -1000.5, 1)) and s["radius"][1] == 1000
    assert np.allclose(s["center1"][7], (-2, 1, -1)) and s["time1"][7] == 1.0
    kinds = [d.materials[int(m)].kind for m in s["material"]]
    assert kinds == [0, 0, 3, 1, 1, 2, 3, 0]
    assert abs(d.camera.focus_dist - np.sqrt(38.0)) < 1e-5 and d.camera.time1 == np.float32(0.2)
    assert d.materials[int(s["material"][3])].param == 0.0 and d.materials[int(s["material"][2])].param == 2.0
    c2 = rt.SceneDesc.builtin("book1_final")
    assert 470 <= c2.desc.n_spheres <= 490 and c2.desc.n_images == 0
    c3 = rt.SceneDesc.builtin("perlin_motion")
    assert c3.desc.n_spheres == 145
    k3 = {c3.desc.textures[i].kind for i in range(c3.desc.n_textures)}
    assert k3 == {capi.RT_TEX_CONSTANT, capi.RT_TEX_CHECKER, capi.RT_TEX_NOISE_PERLIN, capi.RT_TEX_NOISE_TURBULANCE,
                  capi.RT_TEX_NOISE_MARBLE, capi.RT_TEX_WOOD}
    assert (c3.spheres()["flags"] & capi.RT_SPHERE_MOVING).sum() == 18
    c4 = rt.SceneDesc.builtin("random_spheres", n=1000)
    assert c4.desc.n_spheres == 1001 and c4.desc.n_materials <= 513   # shared materials are de-duplicated
    with pytest.raises(capi.RtError):
        rt.SceneDesc.builtin("no_such_scene")
    with pytest.raises(capi.RtError):
        rt.SceneDesc.builtin("earth_emitter")  # needs the image


def test_scene_file_round_trip(tmp_path, earth):
    for name, kw in (("earth_emitter", dict(image=earth)), ("perlin_motion", {})):
        a = rt.SceneDesc.builtin(name, **kw)
        p = tmp_path / f"{name}.rtsc"
        a.save(str(p))
        b = rt.SceneDesc.load(str(p))
        assert a.spheres().tobytes() == b.spheres().tobytes()
        assert bytes(a.desc.camera) == bytes(b.desc.camera)
        assert a.desc.n_textures == b.desc.n_textures and a.desc.n_materials == b.desc.n_materials
        if a.desc.n_images:
            n = a.desc.images[0].width * a.desc.images[0].height * 3
            assert np.array_equal(np.ctypeslib.as_array(a.desc.images[0].rgb, (n,)), np.ctypeslib.as_array(b.desc.images[0].rgb, (n,)))
    with pytest.raises(capi.RtError):
        rt.SceneDesc.load(str(tmp_path / "missing.rtsc"))


def test_damaged_scene_files_are_io_errors(tmp_path, earth):
    """Counts and image sizes of a scene file are checked against the bytes that are there before anything is allocated:
    a header that announces 4 G spheres, an image of 2^31 x 2^31 texels or a truncated file is RT_ERR_IO."""
    import struct

    a = rt.SceneDesc.builtin("earth_emitter", image=earth[:8, :16].copy())
    p = tmp_path / "ok.rtsc"
    a.save(str(p))
    good = p.read_bytes()
    assert rt.SceneDesc.load(str(p)).desc.n_spheres == a.desc.n_spheres
    cam = C.sizeof(capi.rt_camera)
    img_at = 28 + cam + a.desc.n_spheres * C.sizeof(capi.rt_sphere) + a.desc.n_materials * C.sizeof(capi.rt_material) \
        + a.desc.n_textures * C.sizeof(capi.rt_texture)
    assert struct.unpack_from("<ii", good, img_at) == (16, 8)
    cases = {
        "spheres": good[:8] + struct.pack("<I", 0xFFFFFFFF) + good[12:],
        "materials": good[:12] + struct.pack("<I", 0xFFFFFFF0) + good[16:],
        "images": good[:20] + struct.pack("<I", 0x7FFFFFFF) + good[24:],
        "image_size": good[:img_at] + struct.pack("<ii", 0x7FFFFFFF, 0x7FFFFFFF) + good[img_at + 8:],
        "truncated_arrays": good[:28 + cam + 40],
        "truncated_image": good[:-5],
        "header_only": good[:28],
        "empty": b"",
        "magic": b"XXXX" + good[4:],
    }
    for name, blob in cases.items():
        q = tmp_path / f"{name}.rtsc"
        q.write_bytes(blob)
        with pytest.raises(capi.RtError) as e:
            rt.SceneDesc.load(str(q))
        assert e.value.status == capi.RT_ERR_IO, name


def test_writer_conversion_restates_main_cu():
    rng = np.random.default_rng(0)
    rgb = rng.random((7, 5, 3), dtype=np.float32)
    rgb[0, 0] = (1.0, 0.0, 0.999999)
    got = rt.quantize_rgb8(rgb)
    want = ((255.999 * rgb.astype(np.float32)).astype(np.float32).astype(np.int32) & 255).astype(np.uint8)[::-1]
    assert np.array_equal(got, want)  # Y flip + int(255.999f*c) & 255 (main.cu:476-487)


def test_ppm_round_trip(tmp_path):
    lib = capi.load_library()
    img = np.arange(4 * 3 * 3, dtype=np.uint8).reshape(3, 4, 3)
    path = str(tmp_path / "x.ppm").encode()
    assert lib.rt_write_ppm(path, 4, 3, img.ctypes.data) == 0
    out, w, h = C.POINTER(C.c_float)(), C.c_int32(), C.c_int32()
    assert lib.rt_read_ppm_f32(path, C.byref(out), C.byref(w), C.byref(h)) == 0
    got = np.ctypeslib.as_array(out, (3, 4, 3)).copy()
    lib.rt_free(out)
    assert (w.value, h.value) == (4, 3) and np.array_equal(got, img.astype(np.float32) / np.float32(255))


def test_damaged_pnm_files_are_io_errors(tmp_path):
    """The size a PNM header announces must be backed by the file before it sizes an allocation."""
    lib = capi.load_library()
    out, w, h = C.POINTER(C.c_float)(), C.c_int32(), C.c_int32()
    cases = [b"P6\n2000000000 2000000000\n255\n" + b"\0" * 64, b"P5\n65535 65535\n255\nxx", b"P6\n4 3\n255\n" + b"\0" * 35,
             b"P6\n4 3\n65535\n" + b"\0" * 72, b"P6\n-4 3\n255\n" + b"\0" * 36, b"P6", b"", b"P3\n1 1\n255\n0 0 0\n"]
    for i, blob in enumerate(cases):
        p = tmp_path / f"bad{i}.ppm"
        p.write_bytes(blob)
        assert lib.rt_read_ppm_f32(str(p).encode(), C.byref(out), C.byref(w), C.byref(h)) == capi.RT_ERR_IO, blob[:24]
    g = tmp_path / "grey.pgm"
    g.write_bytes(b"P5\n# a comment\n3 2\n255\n" + bytes(range(0, 60, 10)))
    assert lib.rt_read_ppm_f32(str(g).encode(), C.byref(out), C.byref(w), C.byref(h)) == 0
    got = np.ctypeslib.as_array(out, (2, 3, 3)).copy()
    lib.rt_free(out)
    want = (np.arange(0, 60, 10, dtype=np.float32) / np.float32(255)).reshape(2, 3, 1).repeat(3, axis=2)
    assert np.array_equal(got, want)


def test_fastdiv_is_exact(tmp_path):
    """rtd::FastDiv (csrc/rt_fastdiv.hpp) turns the per-path `path % pixels`, `pixel / width` of the kernels into five
    integer instructions; the quotient must be exact for every 32-bit operand (edge values + 8.6 M random pairs)."""
    import subprocess

    exe = tmp_path / "fastdiv_check"
    subprocess.check_call(["g++", "-std=c++17", "-O2", str(ROOT / "tests" / "fastdiv_check.cpp"), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout[-300:]
