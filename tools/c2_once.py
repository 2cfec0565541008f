"""Renders C2 (book-1 final, 485 spheres, host SAH BVH) at 1920x1080x32; used under ncu."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import raytracing_renderer_cuda_b200 as rt
ctx = rt.Context(0)
sc = rt.Scene(ctx, rt.SceneDesc.builtin(sys.argv[1] if len(sys.argv) > 1 else "book1_final"))
for _ in range(2):
    img, st = sc.render(rt.default_params(width=1920, height=1080, spp=32))
print("ms", st.ms_total, "Mrays/s", st.rays / st.ms_total / 1e3, "iterations", st.iterations)
