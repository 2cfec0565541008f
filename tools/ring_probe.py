"""FIRST thing to run on a GPU for the experimental barrier-free kernel (k_wf_ring, RT_WF_GRAIN=ring):

    timeout 300 python tools/ring_probe.py            # small frames, then C1

Every wavefront kernel traces identical paths (the RNG is keyed on pixel/sample/bounce), so the ring kernel must
reproduce the default kernel's frame: equal sample counts per pixel, equal ray counts, colour sums equal up to the
order of the float atomics.  Then the A/B timing on C1 / C2.  Run under `timeout`: a protocol bug shows as a hang."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
import raytracing_renderer_cuda_b200 as rt  # noqa: E402
from raytracing_renderer_cuda_b200.assets import load_earth  # noqa: E402

ctx = rt.Context(0)


def frame(scene, grain, **kw):
    if grain:
        os.environ["RT_WF_GRAIN"] = grain
    else:
        os.environ.pop("RT_WF_GRAIN", None)
    t0 = time.perf_counter()
    acc, st = scene.render_accum(rt.default_params(**kw))
    return acc, st, (time.perf_counter() - t0) * 1e3


cases = [("earth_emitter", dict(width=96, height=48, spp=8)), ("earth_emitter", dict(width=400, height=200, spp=16)),
         ("earth_emitter", dict(width=1200, height=600, spp=100)), ("perlin_motion", dict(width=320, height=160, spp=16)),
         ("book1_final", dict(width=320, height=180, spp=16)), ("book1_final", dict(width=1920, height=1080, spp=64))]
for name, kw in cases:
    sc = rt.Scene(ctx, rt.SceneDesc.builtin(name, load_earth() if name == "earth_emitter" else None))
    ref, sr, _ = frame(sc, None, **kw)
    for rep in range(3):  # the ring and its counters live across frames
        got, st, _ = frame(sc, "ring", **kw)
        d = np.abs(got[..., :3] - ref[..., :3]).max(axis=2) / kw["spp"]
        ok = np.array_equal(got[..., 3], ref[..., 3]) and st.rays == sr.rays and float(d.max()) < 1e-4
        print(name, kw, "rep", rep, "OK" if ok else "MISMATCH", "launches", st.launches, "rays", st.rays, "ref rays", sr.rays,
              "max px diff", float(d.max()), flush=True)
        if not ok:
            sys.exit(1)
    # RT_WF_TAIL (never run on a GPU when this was written): per-iteration kernels for the bulk, one ring launch for the tail
    for tail in ("65536", "1000000"):
        os.environ["RT_WF_TAIL"] = tail
        got, st, _ = frame(sc, None, **kw)
        os.environ.pop("RT_WF_TAIL")
        d = np.abs(got[..., :3] - ref[..., :3]).max(axis=2) / kw["spp"]
        ok = np.array_equal(got[..., 3], ref[..., 3]) and st.rays == sr.rays and float(d.max()) < 1e-4
        print(name, kw, "RT_WF_TAIL", tail, "OK" if ok else "MISMATCH", "launches", st.launches, "iterations", st.iterations,
              "(default:", sr.launches, sr.iterations, ") ms", round(st.ms_total, 3), "default ms", round(sr.ms_total, 3), flush=True)
        if not ok:
            sys.exit(1)
    t = {}
    for grain in (None, "ring", None, "ring"):
        ms = []
        for _ in range(5):
            _, st, _ = frame(sc, grain, **kw)
            ms.append(st.ms_total)
        t.setdefault(grain or "default", []).append(float(np.median(ms)))
    os.environ["RT_WF_TAIL"] = "262144"
    ms = []
    for _ in range(5):
        _, st, _ = frame(sc, None, **kw)
        ms.append(st.ms_total)
    os.environ.pop("RT_WF_TAIL")
    t["default + RT_WF_TAIL=262144"] = [float(np.median(ms))]
    print(name, kw, "ms per frame (median of 5, twice):", t, flush=True)
    sc.close()
os.environ.pop("RT_WF_GRAIN", None)
print("ring probe ok")
