"""ctypes bindings of the checkers: oracle/liboracle.so (this repo's CPU restatement) and, when it
has been built from /root/reference, oracle/_ref/libref_cpu.so (the reference's own headers on the
host).  TEST INFRASTRUCTURE — imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs only."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from raytracing_renderer_cuda_b200 import capi

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "liboracle.so"
REFCPU_SO = ROOT / "oracle" / "_ref" / "libref_cpu.so"
REFCPU_RN_SO = ROOT / "oracle" / "_ref" / "libref_cpu_rn.so"
REFSTB_SO = ROOT / "oracle" / "_ref" / "libref_stb.so"
_VP = C.c_void_p
_F3 = C.c_float * 3


def _f3(p):
    return _F3(*[float(x) for x in p])


class Oracle:
    """oracle/rt_oracle.cpp.  arith: 0 host, 1 reference-GPU contraction; sampler: 0 reference, 1 product."""

    def __init__(self, path: Path = ORACLE_SO):
        if not Path(path).exists():
            raise FileNotFoundError(f"{path}: build it with `make -C oracle liboracle.so`")
        L = self.L = C.CDLL(str(path))
        L.orc_scene_create.restype = _VP
        L.orc_scene_create.argtypes = [C.POINTER(capi.rt_scene_desc)]
        L.orc_scene_destroy.argtypes = [_VP]
        L.orc_trace.argtypes = [_VP, _VP, C.c_size_t, C.c_float, C.c_int, _VP]
        L.orc_render.argtypes = [_VP, C.POINTER(capi.rt_render_params), C.c_int, C.c_int, C.c_int, _VP,
                                 C.POINTER(C.c_ulonglong)]
        L.orc_tonemap.argtypes = [_VP, C.c_int, C.c_int, _VP]
        L.orc_shade_probe.argtypes = [_VP, _VP, C.c_size_t, C.POINTER(capi.rt_render_params), C.c_int, _VP]
        L.orc_perlin_noise.restype = C.c_float
        L.orc_perlin_noise.argtypes = [_F3]
        L.orc_turbulence.restype = C.c_float
        L.orc_turbulence.argtypes = [_F3]
        L.orc_texture_value.argtypes = [_VP, C.c_int, C.c_float, C.c_float, _F3, _F3]
        L.orc_reflect.argtypes = [_F3, _F3, _F3]
        L.orc_refract.restype = C.c_int
        L.orc_refract.argtypes = [_F3, _F3, C.c_float, _F3]
        L.orc_shlick.restype = C.c_float
        L.orc_shlick.argtypes = [C.c_float, C.c_float]
        L.orc_sphere_uv.argtypes = [_F3, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_camera_ray.argtypes = [_VP, C.c_float, C.c_float, C.c_ulonglong, _VP]
        L.orc_light_sample.restype = C.c_float
        L.orc_light_sample.argtypes = [_VP, _F3, _F3, C.c_float, C.c_uint32, C.c_uint32, _F3]

    def scene(self, desc: capi.SceneDesc) -> "OracleScene":
        return OracleScene(self, desc)

    def perlin_noise(self, p):
        return float(self.L.orc_perlin_noise(_f3(p)))

    def turbulence(self, p):
        return float(self.L.orc_turbulence(_f3(p)))

    def reflect(self, v, n):
        out = _F3()
        self.L.orc_reflect(_f3(v), _f3(n), out)
        return np.array(out[:], np.float32)

    def refract(self, v, n, mu):
        out = _F3()
        ok = self.L.orc_refract(_f3(v), _f3(n), mu, out)
        return bool(ok), np.array(out[:], np.float32)

    def shlick(self, c, ri):
        return float(self.L.orc_shlick(c, ri))

    def sphere_uv(self, n):
        u, v = C.c_float(), C.c_float()
        self.L.orc_sphere_uv(_f3(n), C.byref(u), C.byref(v))
        return u.value, v.value

    def tonemap(self, accum: np.ndarray) -> np.ndarray:
        accum = np.ascontiguousarray(accum, np.float32)
        h, w = accum.shape[:2]
        out = np.empty((h, w, 3), np.float32)
        self.L.orc_tonemap(accum.ctypes.data, w, h, out.ctypes.data)
        return out


class OracleScene:
    def __init__(self, o: Oracle, desc: capi.SceneDesc):
        self.o, self.L, self._desc = o, o.L, desc
        self._h = _VP(self.L.orc_scene_create(desc._ptr))

    def trace(self, rays: np.ndarray, tmin: float = 1e-5, arith: int = 1) -> np.ndarray:
        rays = np.ascontiguousarray(rays, capi.RAY_DTYPE)
        hits = np.zeros(len(rays), capi.HIT_DTYPE)
        self.L.orc_trace(self._h, rays.ctypes.data, len(rays), tmin, arith, hits.ctypes.data)
        return hits

    def shade_probe(self, rays: np.ndarray, params: capi.rt_render_params, arith: int = 1) -> np.ndarray:
        rays = np.ascontiguousarray(rays, capi.RAY_DTYPE)
        out = np.zeros(len(rays), capi.SHADE_DTYPE)
        self.L.orc_shade_probe(self._h, rays.ctypes.data, len(rays), C.byref(params), arith, out.ctypes.data)
        return out

    def render(self, params: capi.rt_render_params, sampler: int, arith: int = 0, nthreads: int = 8):
        acc = np.zeros((params.height, params.width, 4), np.float32)
        rays = C.c_ulonglong()
        self.L.orc_render(self._h, C.byref(params), sampler, arith, nthreads, acc.ctypes.data, C.byref(rays))
        return acc, rays.value

    def texture_value(self, tex: int, u: float, v: float, p) -> np.ndarray:
        out = _F3()
        self.L.orc_texture_value(self._h, tex, u, v, _f3(p), out)
        return np.array(out[:], np.float32)

    def light_sample(self, p, n, time: float, seed: int, index: int):
        """RT_RENDER_EMITTER_SAMPLING: the shadow ray of a lambertian hit -> (direction, p_ref / p_sel)."""
        d = _F3()
        w = self.L.orc_light_sample(self._h, _f3(p), _f3(n), time, seed, index, d)
        return np.array(d[:], np.float32), float(w)

    def camera_ray(self, s: float, t: float, seed: int) -> np.ndarray:
        r = np.zeros(1, capi.RAY_DTYPE)
        self.L.orc_camera_ray(self._h, s, t, seed, r.ctypes.data)
        return r

    def __del__(self):
        try:
            self.L.orc_scene_destroy(self._h)
        except Exception:
            pass


class RefCpu:
    """oracle/_ref/libref_cpu.so: the reference headers, unchanged, compiled for the host (oracle/ref_cpu.cpp)."""

    def __init__(self, path: Path = REFCPU_SO):
        if not Path(path).exists():
            raise FileNotFoundError(f"{path}: built only where /root/reference exists (`make -C oracle ref`)")
        L = self.L = C.CDLL(str(path))
        L.refcpu_scene_create.restype = _VP
        L.refcpu_scene_create.argtypes = [C.POINTER(capi.rt_scene_desc), C.c_int]
        L.refcpu_trace.argtypes = [_VP, _VP, C.c_size_t, C.c_float, _VP]
        L.refcpu_render.argtypes = [_VP, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, _VP, _VP, C.POINTER(C.c_ulonglong)]
        L.refcpu_perlin_noise.restype = C.c_float
        L.refcpu_perlin_noise.argtypes = [_F3]
        L.refcpu_turbulence.restype = C.c_float
        L.refcpu_turbulence.argtypes = [_F3]
        L.refcpu_texture_value.argtypes = [_VP, C.c_int, C.c_float, C.c_float, _F3, _F3]
        L.refcpu_reflect.argtypes = [_F3, _F3, _F3]
        L.refcpu_refract.restype = C.c_int
        L.refcpu_refract.argtypes = [_F3, _F3, C.c_float, _F3]
        L.refcpu_shlick.restype = C.c_float
        L.refcpu_shlick.argtypes = [C.c_float, C.c_float]
        L.refcpu_camera_ray.argtypes = [_VP, C.c_float, C.c_float, C.c_ulonglong, _VP]

    def scene(self, desc: capi.SceneDesc, use_bvh: bool = False) -> "RefCpuScene":
        return RefCpuScene(self, desc, use_bvh)

    def perlin_noise(self, p):
        return float(self.L.refcpu_perlin_noise(_f3(p)))

    def turbulence(self, p):
        return float(self.L.refcpu_turbulence(_f3(p)))

    def reflect(self, v, n):
        out = _F3()
        self.L.refcpu_reflect(_f3(v), _f3(n), out)
        return np.array(out[:], np.float32)

    def refract(self, v, n, mu):
        out = _F3()
        ok = self.L.refcpu_refract(_f3(v), _f3(n), mu, out)
        return bool(ok), np.array(out[:], np.float32)

    def shlick(self, c, ri):
        return float(self.L.refcpu_shlick(c, ri))


class RefCpuScene:
    def __init__(self, r: RefCpu, desc: capi.SceneDesc, use_bvh: bool):
        self.L, self._desc = r.L, desc
        self._h = _VP(self.L.refcpu_scene_create(desc._ptr, int(use_bvh)))

    def trace(self, rays: np.ndarray, tmin: float = 1e-5) -> np.ndarray:
        rays = np.ascontiguousarray(rays, capi.RAY_DTYPE)
        hits = np.zeros(len(rays), capi.HIT_DTYPE)
        self.L.refcpu_trace(self._h, rays.ctypes.data, len(rays), tmin, hits.ctypes.data)
        return hits

    def render(self, width: int, height: int, spp: int, seed: int = 1000, nthreads: int = 8, want_fb: bool = True):
        mean = np.zeros((height, width, 3), np.float32)
        fb = np.zeros((height, width, 3), np.float32) if want_fb else None
        rays = C.c_ulonglong()
        self.L.refcpu_render(self._h, width, height, spp, seed, nthreads, mean.ctypes.data,
                             fb.ctypes.data if fb is not None else None, C.byref(rays))
        return mean, fb, rays.value

    def texture_value(self, tex: int, u: float, v: float, p) -> np.ndarray:
        out = _F3()
        self.L.refcpu_texture_value(self._h, tex, u, v, _f3(p), out)
        return np.array(out[:], np.float32)

    def camera_ray(self, s: float, t: float, seed: int) -> np.ndarray:
        r = np.zeros(1, capi.RAY_DTYPE)
        self.L.refcpu_camera_ray(self._h, s, t, seed, r.ctypes.data)
        return r


class FileDesc:
    """A flat scene file (rt_scene_desc_save, include/rt_api.h) read WITHOUT the product library: pure ctypes / numpy.
    Quacks like capi.SceneDesc for the checkers (`._ptr`, `.desc`).  Used by bench.py's reference arm, whose process must
    not load librt_b200.so."""

    def __init__(self, path):
        raw = Path(path).read_bytes()
        hdr = np.frombuffer(raw, np.uint32, 7)
        if hdr[0] != 0x43535452 or hdr[1] != 1:
            raise ValueError(f"{path}: not a scene file")
        n_s, n_m, n_t, n_i, bvh = (int(x) for x in hdr[2:7])
        off = 28
        d = capi.rt_scene_desc()
        C.memmove(C.byref(d.camera), raw[off:off + C.sizeof(capi.rt_camera)], C.sizeof(capi.rt_camera))
        off += C.sizeof(capi.rt_camera)
        self._keep = []

        def take(T, n):
            nonlocal off
            arr = (T * max(n, 1))()
            C.memmove(arr, raw[off:off + n * C.sizeof(T)], n * C.sizeof(T))
            off += n * C.sizeof(T)
            self._keep.append(arr)
            return arr

        d.spheres, d.n_spheres = take(capi.rt_sphere, n_s), n_s
        d.materials, d.n_materials = take(capi.rt_material, n_m), n_m
        d.textures, d.n_textures = take(capi.rt_texture, n_t), n_t
        imgs = (capi.rt_image * max(n_i, 1))()
        for i in range(n_i):
            w, h = (int(x) for x in np.frombuffer(raw, np.int32, 2, off))
            off += 8
            px = np.frombuffer(raw, np.float32, w * h * 3, off).copy()
            off += px.nbytes
            self._keep.append(px)
            imgs[i].rgb = px.ctypes.data_as(C.POINTER(C.c_float))
            imgs[i].width, imgs[i].height = w, h
        self._keep.append(imgs)
        d.images, d.n_images, d.bvh_mode = imgs, n_i, bvh
        self.desc = d
        self._ptr = C.pointer(d)


def camera_rays(desc: capi.SceneDesc, n: int, seed: int = 1) -> np.ndarray:
    """Pin-hole rays through random image positions of the scene camera with random shutter times
    (host float math; test INPUT only — both sides of every comparison receive the same rays)."""
    c = desc.desc.camera
    rng = np.random.default_rng(seed)
    lf, la, up = (np.array(v[:], np.float32) for v in (c.lookfrom, c.lookat, c.up))
    hh = np.tan(np.float32(c.vfov * np.pi / 180.0) / 2)
    hw = c.aspect * hh
    wv = (lf - la) / np.linalg.norm(lf - la)
    u = np.cross(up, wv)
    u /= np.linalg.norm(u)
    v = np.cross(wv, u)
    fd = c.focus_dist
    ll = lf - hw * fd * u - hh * fd * v - fd * wv
    s = rng.random(n, dtype=np.float32)[:, None]
    t = rng.random(n, dtype=np.float32)[:, None]
    rays = np.zeros(n, dtype=capi.RAY_DTYPE)
    rays["origin"] = lf
    rays["direction"] = (ll + s * (2 * hw * fd * u) + t * (2 * hh * fd * v) - lf).astype(np.float32)
    rays["time"] = (c.time0 + rng.random(n, dtype=np.float32) * (c.time1 - c.time0)).astype(np.float32)
    return rays


def secondary_rays(desc: capi.SceneDesc, hits: np.ndarray, seed: int = 2) -> np.ndarray:
    """Rays leaving the surface points of `hits` in random directions: exercises tmin self-intersection,
    inside-sphere roots and rays that start on the r = 1000 ground."""
    rng = np.random.default_rng(seed)
    ok = hits["id"] != capi.RT_INVALID_ID
    h = hits[ok]
    d = rng.normal(size=(len(h), 3)).astype(np.float32)
    d = d + h["n"]
    rays = np.zeros(len(h), dtype=capi.RAY_DTYPE)
    rays["origin"] = h["p"]
    rays["direction"] = d
    c = desc.desc.camera
    rays["time"] = (c.time0 + rng.random(len(h), dtype=np.float32) * (c.time1 - c.time0)).astype(np.float32)
    return rays


def _jpeg_call(fn, rgb8: np.ndarray, quality: int) -> bytes:
    rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
    h, w = rgb8.shape[:2]
    cap = 2048 + ((w + 15) // 8) * ((h + 15) // 8) * 3 * 440
    out = np.empty(cap, dtype=np.uint8)
    n = fn(rgb8.ctypes.data, w, h, quality, out.ctypes.data, cap)
    assert n > 0, "encoder returned 0 bytes"
    return out[:n].tobytes()


def oracle_jpeg(rgb8: np.ndarray, quality: int = 100) -> bytes:
    """oracle/jpeg_oracle.cpp: CPU restatement of stbi_write_jpg (main.cu:491)."""
    L = C.CDLL(str(ORACLE_SO))
    L.orc_jpeg_encode.restype = C.c_size_t
    L.orc_jpeg_encode.argtypes = [_VP, C.c_int, C.c_int, C.c_int, _VP, C.c_size_t]
    return _jpeg_call(L.orc_jpeg_encode, rgb8, quality)


def ref_stb_jpeg(rgb8: np.ndarray, quality: int = 100) -> bytes:
    """oracle/_ref/libref_stb.so: the reference's vendored stb_image_write.h itself."""
    L = C.CDLL(str(REFSTB_SO))
    L.ref_stb_write_jpg.restype = C.c_size_t
    L.ref_stb_write_jpg.argtypes = [_VP, C.c_int, C.c_int, C.c_int, _VP, C.c_size_t]
    return _jpeg_call(L.ref_stb_write_jpg, rgb8, quality)


def jpeg_test_image(kind: str, w: int, h: int, seed: int = 0) -> np.ndarray:
    """Deterministic rgb8 inputs for the JPEG parity tests (numpy Generator streams are stable across versions)."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elif kind == "smooth":
        y, x = np.mgrid[0:h, 0:w]
        img = np.stack([x * 255 / max(w - 1, 1), y * 255 / max(h - 1, 1), (x + y) * 127 / max(w + h - 2, 1)], -1).astype(np.uint8)
    elif kind == "flat":
        img = np.full((h, w, 3), 200, np.uint8)
    elif kind == "sat":  # saturated checker noise: many 0xFF bytes in the stream, long zero runs at low quality
        img = (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)
    elif kind == "photo":  # smooth + mild noise, like a rendered frame
        y, x = np.mgrid[0:h, 0:w]
        base = 128 + 90 * np.sin(x / 17.0)[..., None] * np.cos(y / 11.0)[..., None] * np.array([1.0, 0.7, 0.4])
        img = np.clip(base + rng.normal(0, 6, (h, w, 3)), 0, 255).astype(np.uint8)
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(img)


def oracle_jpeg_pixels(coeffs_ptr) -> np.ndarray:
    """oracle/jpeg_decode_oracle.cpp: coefficient planes (capi.jpeg_parse) -> [H, W, ch] bytes."""
    L = C.CDLL(str(ORACLE_SO))
    L.orc_jpeg_pixels.restype = C.c_int
    L.orc_jpeg_pixels.argtypes = [_VP, _VP]
    c = coeffs_ptr.contents
    ch = 1 if c.n_comp == 1 else 3
    out = np.zeros((c.height, c.width, ch), dtype=np.uint8)
    assert L.orc_jpeg_pixels(C.cast(coeffs_ptr, _VP), out.ctypes.data) == 0
    return out


def ref_stb_load_jpeg(file_bytes: bytes) -> np.ndarray:
    """oracle/_ref/libref_stb.so: stbi_load_from_memory(..., 0) of the reference's vendored stb_image.h -> [H, W, ch] bytes."""
    L = C.CDLL(str(REFSTB_SO))
    L.ref_stb_load_jpg.restype = C.c_int
    L.ref_stb_load_jpg.argtypes = [_VP, C.c_int, _VP, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    buf = np.frombuffer(file_bytes, dtype=np.uint8)
    cap = 64 << 20
    out = np.zeros(cap, dtype=np.uint8)
    w, h = C.c_int(), C.c_int()
    ch = L.ref_stb_load_jpg(buf.ctypes.data, buf.size, out.ctypes.data, cap, C.byref(w), C.byref(h))
    assert ch > 0, "stb could not decode the file"
    return out[:w.value * h.value * ch].reshape(h.value, w.value, ch).copy()
