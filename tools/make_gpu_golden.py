"""Records golden outputs of the REFERENCE's own CUDA kernels on a B200 (oracle/_ref/ref_harness, which
compiles the reference translation unit unchanged): closest hits for the rays of
tests/golden/ref_cpu_golden.npz and converged 4096-spp renders.  Run on the GPU box:

    gpurun -- python tools/make_gpu_golden.py      # writes gpurun_out/golden/ref_gpu_golden.npz

then copy the file to tests/golden/.  Scratch files go to /tmp (gpurun_out/ is size-capped)."""
import json
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import raytracing_renderer_cuda_b200 as rt  # noqa: E402
from raytracing_renderer_cuda_b200 import capi  # noqa: E402
from raytracing_renderer_cuda_b200.assets import load_earth  # noqa: E402

HARNESS = ROOT / "oracle" / "_ref" / "ref_harness"
tmp = Path(tempfile.mkdtemp(prefix="rtgold"))
cpu = np.load(ROOT / "tests" / "golden" / "ref_cpu_golden.npz")
out = {}
meta = {}
RENDERS = {"earth_emitter": (400, 200, 4096), "book1_final": (320, 180, 4096), "perlin_motion": (300, 150, 4096)}
earth = load_earth()
for name, kw in (("earth_emitter", dict(image=earth)), ("book1_final", {}), ("perlin_motion", {})):
    d = rt.SceneDesc.builtin(name, **kw)
    sp = tmp / f"{name}.rtsc"
    d.save(str(sp))
    rays = np.ascontiguousarray(cpu[f"{name}_rays"]).view(capi.RAY_DTYPE).reshape(-1)
    for use_bvh in (1, 0):
        rays.tofile(tmp / "rays.bin")
        subprocess.check_call([str(HARNESS), "trace", str(sp), str(tmp / "rays.bin"), str(tmp / "hits.bin"), str(use_bvh)])
        out[f"{name}_hits_bvh{use_bvh}"] = np.fromfile(tmp / "hits.bin", dtype=capi.HIT_DTYPE)
    w, h, spp = RENDERS[name]
    o = subprocess.check_output([str(HARNESS), "render", str(sp), str(w), str(h), str(spp), str(tmp / "r.f32"), "1", "1", "1"],
                                text=True)
    info = json.loads(o.strip().splitlines()[-1])
    a = np.fromfile(tmp / "r.f32", dtype=np.float32).reshape(2, h, w, 3)
    out[f"{name}_fb_{w}x{h}x{spp}"] = a[0]
    out[f"{name}_mean_{w}x{h}x{spp}"] = a[1].astype(np.float16)  # un-tonemapped mean, informational
    meta[name] = info
    print(name, info, flush=True)
dst = ROOT / "gpurun_out" / "golden"
dst.mkdir(parents=True, exist_ok=True)
np.savez_compressed(dst / "ref_gpu_golden.npz", **out)
json.dump(meta, open(dst / "ref_gpu_golden.json", "w"), indent=1)
print("wrote", dst / "ref_gpu_golden.npz", (dst / "ref_gpu_golden.npz").stat().st_size)
