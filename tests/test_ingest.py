"""Image-texture ingest and the reference's second scene, CPU side: channel conversion, PGM/PPM
reading, and the structure of the hdr_sphere restatement (populate_scene_hdr, main.cu:136-182)."""
import ctypes as C

import numpy as np
import pytest

import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi


def test_channel_counts_follow_stb_conversion():
    rng = np.random.default_rng(0)
    g = rng.random((5, 7, 1), dtype=np.float32)
    assert np.array_equal(capi.image_to_rgb(g), np.repeat(g, 3, axis=2))
    ga = rng.random((5, 7, 2), dtype=np.float32)
    assert np.array_equal(capi.image_to_rgb(ga), np.repeat(ga[..., :1], 3, axis=2))
    rgb = rng.random((5, 7, 3), dtype=np.float32)
    assert np.array_equal(capi.image_to_rgb(rgb), rgb)
    rgba = rng.random((5, 7, 4), dtype=np.float32)
    assert np.array_equal(capi.image_to_rgb(rgba), rgba[..., :3])
    lib = capi.load_library()
    assert lib.rt_image_to_rgb(rgb.ctypes.data, 7, 5, 5, rgb.ctypes.data) == capi.RT_ERR_INVALID_ARG


def test_pgm_and_ppm_files_load_as_rgb(tmp_path):
    lib = capi.load_library()
    rng = np.random.default_rng(1)
    for magic, ch in ((b"P6", 3), (b"P5", 1)):
        b = rng.integers(0, 256, (9, 13, ch), dtype=np.uint8)
        f = tmp_path / f"img{ch}.pnm"
        f.write_bytes(magic + b"\n# comment\n13 9\n255\n" + b.tobytes())
        ptr, w, h = C.POINTER(C.c_float)(), C.c_int32(), C.c_int32()
        assert lib.rt_read_ppm_f32(str(f).encode(), C.byref(ptr), C.byref(w), C.byref(h)) == capi.RT_OK
        got = np.ctypeslib.as_array(ptr, (9, 13, 3)).copy()
        lib.rt_free(ptr)
        want = np.repeat(b, 3 // ch, axis=2).astype(np.float32) / np.float32(255)  # stbi_loadf: byte / 255.f (main.cu:378-380)
        assert (w.value, h.value) == (13, 9) and np.array_equal(got, want)


def test_hdr_sphere_is_the_reference_scene(scene_descs):
    d = scene_descs["hdr_sphere"]
    desc = d.desc
    assert desc.n_spheres == 3 and desc.n_images == 1 and desc.bvh_mode == capi.RT_BVH_NONE  # hitable_list(objects, nullptr, 3)
    sp = d.spheres()
    assert list(sp["id"]) == [0, 1, 2] and list(sp["radius"]) == [1.0, 10.0, 1.0]
    assert np.array_equal(sp["center0"], np.array([[1, 0, -1], [0, 0, 0], [-1, 0, -1]], np.float32))
    kinds = [desc.materials[int(m)].kind for m in sp["material"]]
    assert kinds == [capi.RT_MAT_METAL, capi.RT_MAT_EMITTER, capi.RT_MAT_LAMBERTIAN]
    assert abs(desc.materials[int(sp["material"][0])].param - 0.05) < 1e-7
    c = desc.camera
    assert tuple(c.lookfrom) == (-1.0, 2.0, 9.0) and tuple(c.lookat) == (0.0, 0.0, -1.0)
    assert abs(c.focus_dist - np.float32(np.sqrt(np.float32(105.0)))) < 1e-5 and abs(c.aperture - 0.25) < 1e-7
    assert (c.time0, c.time1) == (0.0, np.float32(0.2))
    assert (desc.images[0].width, desc.images[0].height) == (250, 130)
    with pytest.raises(capi.RtError):
        rt.SceneDesc.builtin("hdr_sphere")  # needs its environment image
