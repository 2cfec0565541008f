"""One full C4 frame (1 M spheres, 3840x2160x64) for a launch list: per-iteration durations of k_wf_step_pt."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import raytracing_renderer_cuda_b200 as rt
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = rt.Context(0)
sc = rt.Scene(ctx, rt.SceneDesc.builtin("random_spheres", n=1_000_000))
img, st = sc.render(rt.default_params(width=3840, height=2160, spp=spp))
print("render ms", st.ms_total, "Mrays/s", st.rays / st.ms_total / 1e3, "iterations", st.iterations, "rays", st.rays, "paths", st.paths)
