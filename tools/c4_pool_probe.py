"""C4 (1M spheres) at 3840x2160x8: path-pool size sweep (RT_WF_POOL)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import raytracing_renderer_cuda_b200 as rt
ctx = rt.Context(0)
sc = rt.Scene(ctx, rt.SceneDesc.builtin("random_spheres", n=1_000_000))
for pool in (1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24):
    os.environ["RT_WF_POOL"] = str(pool)
    best = 1e9
    for _ in range(2):
        img, st = sc.render(rt.default_params(width=3840, height=2160, spp=8))
        best = min(best, st.ms_total)
    print("pool", pool, "ms", round(best, 1), "Mrays/s", round(st.rays / best / 1e3, 1), "iterations", st.iterations, flush=True)
