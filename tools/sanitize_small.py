"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
from raytracing_renderer_cuda_b200.assets import load_earth
from tests import oracle_api as oa
from tests.conftest import make_env_image
ctx = rt.Context(0)
for name, kw in (("earth_emitter", dict(image=load_earth())), ("book1_final", {}), ("perlin_motion", {}), ("hdr_sphere", dict(image=make_env_image(33, 17))),
                 ("random_spheres", dict(n=6000))):
    sc = rt.Scene(ctx, rt.SceneDesc.builtin(name, **kw))
    for grain in ("auto", "pt", "cta"):
        if grain == "auto":
            os.environ.pop("RT_WF_GRAIN", None)
        else:
            os.environ["RT_WF_GRAIN"] = grain
        img, st = sc.render(rt.default_params(width=97, height=41, spp=3))
        img, st = sc.render(rt.default_params(width=61, height=33, spp=2, flags=capi.RT_RENDER_EMITTER_SAMPLING))
    os.environ.pop("RT_WF_GRAIN", None)
    img, st = sc.render(rt.default_params(width=64, height=32, spp=2, pipeline=capi.RT_PIPE_MEGAKERNEL))
    img, st = sc.render(rt.default_params(width=64, height=32, spp=2, pipeline=capi.RT_PIPE_MEGAKERNEL, flags=capi.RT_RENDER_EMITTER_SAMPLING))
    f, st = sc.render_jpeg(rt.default_params(width=97, height=41, spp=2), 100)
    f2, st = sc.render_jpeg(rt.default_params(width=50, height=30, spp=1), 60)
    sc.render_progressive(rt.default_params(width=40, height=20, spp=1), 2)
    print(name, "ok", len(f), len(f2), flush=True)
for (w, h) in ((1, 1), (8, 8), (37, 21), (300, 200)):
    for q in (100, 50):
        img = oa.jpeg_test_image("sat", w, h, 3)
        assert capi.jpeg_encode(ctx, img, q) == oa.oracle_jpeg(img, q)
print("jpeg ok")
