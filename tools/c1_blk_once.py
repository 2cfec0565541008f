"""Renders the C1 frame three times with the record-block kernel (RT_WF_GRAIN=blk); used under ncu."""
import os, sys
from pathlib import Path
os.environ["RT_WF_GRAIN"] = "blk"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from raytracing_renderer_cuda_b200 import capi
if len(sys.argv) > 1:
    capi.LIB_PATH = Path(sys.argv[1])
import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200.assets import load_earth
ctx = rt.Context(0)
sc = rt.Scene(ctx, rt.SceneDesc.builtin("earth_emitter", image=load_earth()))
for _ in range(3):
    img, st = sc.render(rt.default_params())
print("ms", st.ms_total, "iterations", st.iterations, "launches", st.launches)
