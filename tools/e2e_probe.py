import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200.assets import load_earth
desc = rt.SceneDesc.builtin("earth_emitter", load_earth())
ctx = rt.Context(0)
p = rt.default_params()
out = np.empty((600, 1200, 3), np.float32)
for k in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); sc = rt.Scene(ctx, desc); t1 = time.perf_counter()
    img, st = sc.render(p, out); t2 = time.perf_counter()
    sc.close(); t3 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.2f} ms  render {1e3*(t2-t1):.2f} ms (device {st.ms_total:.2f}, d2h {st.ms_d2h:.2f})  destroy {1e3*(t3-t2):.2f} ms", flush=True)
