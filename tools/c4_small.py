import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import raytracing_renderer_cuda_b200 as rt
d = rt.SceneDesc.builtin("random_spheres", n=1_000_000)
ctx = rt.Context(0)
sc = rt.Scene(ctx, d)
img, st = sc.render(rt.default_params(width=1920, height=1080, spp=4))
print("render ms", st.ms_total, "Mrays/s", st.rays / st.ms_total / 1e3, st.iterations)
