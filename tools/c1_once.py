"""Renders the C1 frame (1200x600x100) three times; used under `ncu --metrics gpu__time_duration.sum` for the launch list."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200.assets import load_earth
ctx = rt.Context(0)
sc = rt.Scene(ctx, rt.SceneDesc.builtin("earth_emitter", image=load_earth()))
for _ in range(3):
    img, st = sc.render(rt.default_params())
print("ms", st.ms_total, "iterations", st.iterations, "launches", st.launches)
