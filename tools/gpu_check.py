"""First-light GPU check: renders C1/C2 with both pipelines, times the reference binary and
harness next to them, and compares closest hits with the reference kernel.  Writes
gpurun_out/gpu_check.json.  Run on the GPU box: `python tools/gpu_check.py`."""
import json
import os
import re
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import raytracing_renderer_cuda_b200 as rt  # noqa: E402
from raytracing_renderer_cuda_b200 import capi  # noqa: E402
from raytracing_renderer_cuda_b200.assets import load_earth  # noqa: E402

import tempfile
RES = ROOT / "gpurun_out"
RES.mkdir(exist_ok=True)
OUT = Path(tempfile.mkdtemp(prefix="rtcheck"))  # scratch: gpurun_out/ is size-capped
REF = ROOT / "oracle" / "_ref"
res = {}


def camera_rays(desc, w, h, n, seed=1):
    """pin-hole rays through random pixels of the scene camera (host float math; only used as test input)"""
    c = desc.desc.camera
    rng = np.random.default_rng(seed)
    lf, la, up = (np.array(v[:], np.float32) for v in (c.lookfrom, c.lookat, c.up))
    theta = np.float32(c.vfov * np.pi / 180.0)
    hh = np.tan(theta / 2)
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


def ref_trace(scene_path, rays, use_bvh=1):
    rp, hp = OUT / "rays.bin", OUT / "hits.bin"
    rays.tofile(rp)
    subprocess.check_call([str(REF / "ref_harness"), "trace", str(scene_path), str(rp), str(hp), str(use_bvh)])
    return np.fromfile(hp, dtype=capi.HIT_DTYPE)


def ref_render(scene_path, w, h, spp, use_bvh=1, reps=1):
    op = OUT / "ref_render.f32"
    o = subprocess.check_output([str(REF / "ref_harness"), "render", str(scene_path), str(w), str(h), str(spp), str(op),
                                 str(use_bvh), "1", str(reps)], text=True)
    info = json.loads(o.strip().splitlines()[-1])
    a = np.fromfile(op, dtype=np.float32).reshape(2, h, w, 3)
    return a[0], a[1], info


def compare_hits(mine, ref):
    ids_equal = mine["id"] == ref["id"]
    hit = ref["id"] != capi.RT_INVALID_ID
    both = ids_equal & hit
    rel = np.abs(mine["t"][both] - ref["t"][both]) / np.abs(ref["t"][both])
    return {"n": int(len(ref)), "hits": int(hit.sum()), "id_mismatch": int((~ids_equal).sum()),
            "t_bit_exact": int((mine["t"][both] == ref["t"][both]).sum()), "t_max_rel": float(rel.max() if len(rel) else 0),
            "p_max_abs": float(np.abs(mine["p"][both] - ref["p"][both]).max() if both.any() else 0),
            "n_max_abs": float(np.abs(mine["n"][both] - ref["n"][both]).max() if both.any() else 0),
            "uv_max_abs": float(max(np.abs(mine["u"][both] - ref["u"][both]).max(), np.abs(mine["v"][both] - ref["v"][both]).max()) if both.any() else 0)}


ctx = rt.Context(0)
earth = load_earth()

# ---------------- C1 ----------------
c1 = rt.SceneDesc.builtin("earth_emitter", earth)
c1_path = OUT / "c1.rtsc"
c1.save(str(c1_path))
sc1 = rt.Scene(ctx, c1)
for pipe, name in ((capi.RT_PIPE_WAVEFRONT, "wavefront"), (capi.RT_PIPE_MEGAKERNEL, "mega")):
    p = rt.default_params(pipeline=pipe)
    best = None
    for rep in range(3):
        img, st = sc1.render(p)
        if best is None or st.ms_total < best["ms"]:
            best = {"ms": st.ms_total, "paths": st.paths, "rays": st.rays, "launches": st.launches,
                    "iterations": st.iterations, "mpaths_s": st.paths / st.ms_total / 1e3,
                    "mrays_s": st.rays / st.ms_total / 1e3, "ms_tonemap": st.ms_tonemap, "ms_d2h": st.ms_d2h}
    res[f"c1_{name}"] = best
    np.save(OUT / f"c1_{name}.npy", img)
    print(name, best, flush=True)

# reference binary, unchanged (its own chrono window)
t0 = time.time()
o = subprocess.run([str(REF / "ref_main")], cwd=str(REF), capture_output=True, text=True)
m = re.search(r"took (\d+)us", o.stdout)
res["c1_ref_main"] = {"took_us": int(m.group(1)) if m else None, "wall_s": time.time() - t0, "rc": o.returncode}
print("ref_main", res["c1_ref_main"], o.stdout[-300:], o.stderr[-300:], flush=True)
fb, mean, info = ref_render(c1_path, 1200, 600, 100, use_bvh=1, reps=2)
res["c1_ref_harness"] = info
print("ref_harness", info, flush=True)
mine = np.load(OUT / "c1_wavefront.npy")
res["c1_psnr_100spp_wavefront_vs_ref"] = rt.psnr(mine, fb)
res["c1_psnr_100spp_mega_vs_ref"] = rt.psnr(np.load(OUT / "c1_mega.npy"), fb)
res["c1_psnr_wavefront_vs_mega"] = rt.psnr(mine, np.load(OUT / "c1_mega.npy"))
print({k: v for k, v in res.items() if "psnr" in k}, flush=True)

rays = camera_rays(c1, 1200, 600, 1 << 20)
mh = sc1.trace_primary(rays, use_bvh=False)
rh = ref_trace(c1_path, rays, use_bvh=1)
res["c1_trace_vs_ref_bvh"] = compare_hits(mh, rh)
rh0 = ref_trace(c1_path, rays, use_bvh=0)
res["c1_trace_vs_ref_list"] = compare_hits(mh, rh0)
print(res["c1_trace_vs_ref_bvh"], res["c1_trace_vs_ref_list"], flush=True)

# ---------------- C2 (reduced) ----------------
c2 = rt.SceneDesc.builtin("book1_final")
c2_path = OUT / "c2.rtsc"
c2.save(str(c2_path))
sc2 = rt.Scene(ctx, c2)
i2 = sc2.info()
res["c2_info"] = {"n": i2.n_spheres, "nodes": i2.n_nodes, "mode": i2.bvh_mode, "depth": i2.bvh_depth, "ms_build": i2.ms_build,
                  "sah": i2.sah_cost}
W2, H2, S2 = 960, 540, 32
for pipe, name in ((capi.RT_PIPE_WAVEFRONT, "wavefront"), (capi.RT_PIPE_MEGAKERNEL, "mega")):
    p = rt.default_params(pipeline=pipe, width=W2, height=H2, spp=S2)
    best = None
    for rep in range(2):
        img, st = sc2.render(p)
        if best is None or st.ms_total < best["ms"]:
            best = {"ms": st.ms_total, "paths": st.paths, "rays": st.rays, "launches": st.launches,
                    "iterations": st.iterations, "mpaths_s": st.paths / st.ms_total / 1e3, "mrays_s": st.rays / st.ms_total / 1e3}
    res[f"c2_{name}"] = best
    np.save(OUT / f"c2_{name}.npy", img)
    print("c2", name, best, flush=True)
fb2, mean2, info2 = ref_render(c2_path, W2, H2, S2, use_bvh=1)
res["c2_ref_harness"] = info2
res["c2_psnr_vs_ref"] = rt.psnr(np.load(OUT / "c2_wavefront.npy"), fb2)
print("c2 ref", info2, res["c2_psnr_vs_ref"], flush=True)
rays2 = camera_rays(c2, W2, H2, 1 << 20, seed=2)
m_bvh = sc2.trace_primary(rays2, use_bvh=True)
m_list = sc2.trace_primary(rays2, use_bvh=False)
res["c2_bvh_vs_list_mine"] = compare_hits(m_bvh, m_list)
r2 = ref_trace(c2_path, rays2, use_bvh=1)
res["c2_trace_vs_ref_bvh"] = compare_hits(m_bvh, r2)
print(res["c2_bvh_vs_list_mine"], res["c2_trace_vs_ref_bvh"], flush=True)

json.dump(res, open(RES / "gpu_check.json", "w"), indent=1)
print(json.dumps(res))
