"""How much of the C1 frame is the thin tail of long paths?  Renders with max_depth 50 / 16 / 10 and tail fusion on/off."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from raytracing_renderer_cuda_b200 import capi
if len(sys.argv) > 1:
    capi.LIB_PATH = Path(sys.argv[1])
import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200.assets import load_earth
ctx = rt.Context(0)
sc = rt.Scene(ctx, rt.SceneDesc.builtin("earth_emitter", image=load_earth()))
for tail in ("0", "65536", "262144"):
    os.environ["RT_WF_TAIL"] = tail
    for depth in (50, 16, 10, 6):
        best = 1e9
        for _ in range(5):
            img, st = sc.render(rt.default_params(max_depth=depth))
            best = min(best, st.ms_total)
        print("tail_max", tail, "max_depth", depth, "ms", round(best, 3), "iterations", st.iterations, "rays", st.rays, flush=True)
