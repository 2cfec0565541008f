"""Cost and effect of RT_RENDER_EMITTER_SAMPLING on the device (run through gpurun; prints one JSON line per case).
C1 (sky as bright as the emitters: the flag only costs) and the lamp-lit room of tests/test_emitter_sampling.py
(dark world: the flag is what makes the frame converge)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import raytracing_renderer_cuda_b200 as rt  # noqa: E402
from raytracing_renderer_cuda_b200 import capi  # noqa: E402
from raytracing_renderer_cuda_b200.assets import load_earth  # noqa: E402
from tests.test_emitter_sampling import DARK, lit_room_desc  # noqa: E402

NEE = capi.RT_RENDER_EMITTER_SAMPLING


def timed(sc, **kw):
    best = None
    for _ in range(3):
        acc, st = sc.render_accum(rt.default_params(**kw))
        if best is None or st.ms_total < best[1].ms_total:
            best = (acc, st)
    return best


def main():
    ctx = rt.Context(0)
    c1 = rt.Scene(ctx, rt.SceneDesc.builtin("earth_emitter", load_earth()))
    room = rt.Scene(ctx, lit_room_desc())
    for name, sc, kw in (("c1", c1, dict(width=1200, height=600, spp=100)),
                         ("lit_room", room, dict(width=1200, height=600, spp=100, max_depth=8, **DARK))):
        out = {"scene": name}
        lit = None
        for tag, flags in (("nee", NEE), ("plain", 0)):
            a, st = timed(sc, flags=flags, seed=1, **kw)
            b, _ = sc.render_accum(rt.default_params(flags=flags, seed=2, **kw))
            fa, fb = a[..., :3].astype(np.float64) / kw["spp"], b[..., :3].astype(np.float64) / kw["spp"]
            if lit is None:  # pixels that do not see a lamp, decided on the low-noise frames
                lit = np.maximum(fa, fb).max(axis=2) < 0.6 if name == "lit_room" else np.ones(fa.shape[:2], bool)
            out[tag] = {"ms": round(st.ms_total, 3), "rays": int(st.rays), "mrays_s": round(st.rays / st.ms_total / 1e3, 1),
                        "mean": float(fa[lit].mean()), "var_per_pixel_at_spp": float(np.mean((fa - fb)[lit] ** 2) / 2)}
        p, n = out["plain"], out["nee"]
        # time to reach the same noise level: variance x frame time
        out["nee_efficiency_gain"] = round((p["var_per_pixel_at_spp"] * p["ms"]) / (n["var_per_pixel_at_spp"] * n["ms"]), 2)
        print(json.dumps(out))


if __name__ == "__main__":
    main()
