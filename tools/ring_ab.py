"""A/B of k_wf_ring build variants on C1 (tools/ring_first_job.sh / by hand): one process per library, RT_WF_GRAIN=ring,
1 warm-up + 5 frames, median.  `python tools/ring_ab.py ring_b4 ring_e ring_r ring_s ring_all default`"""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
if len(sys.argv) > 2 and sys.argv[1] == "--child":
    sys.path.insert(0, str(ROOT))
    import numpy as np
    from raytracing_renderer_cuda_b200 import capi
    if sys.argv[2] != "default":
        capi.LIB_PATH = ROOT / "gpurun_variants" / f"librt_{sys.argv[2]}.so"
    import raytracing_renderer_cuda_b200 as rt
    from raytracing_renderer_cuda_b200.assets import load_earth
    ctx = rt.Context(0)
    sc = rt.Scene(ctx, rt.SceneDesc.builtin("earth_emitter", image=load_earth()))
    ms = []
    for k in range(6):
        _, st = sc.render_accum(rt.default_params(width=1200, height=600, spp=100))
        ms.append(st.ms_total)
    print(sys.argv[2], "C1 ms/frame median", round(float(np.median(ms[1:])), 3), "rays", st.rays, "launches", st.launches, flush=True)
else:
    os.environ["RT_WF_GRAIN"] = "ring"
    for v in sys.argv[1:]:
        o = subprocess.run([sys.executable, __file__, "--child", v], capture_output=True, text=True, timeout=30)
        print(o.stdout.strip() or o.stderr[-300:], flush=True)
