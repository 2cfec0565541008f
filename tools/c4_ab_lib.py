"""C4 probe/full frame of library variants, each in its own process:  python tools/c4_ab_lib.py [full] name[@ENV=v,...] ..."""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, str(ROOT))
    from raytracing_renderer_cuda_b200 import capi
    capi.LIB_PATH = Path(sys.argv[2])
    import raytracing_renderer_cuda_b200 as rt
    ctx = rt.Context(0)
    sc = rt.Scene(ctx, rt.SceneDesc.builtin("random_spheres", n=1_000_000))
    out = []
    for w, h, spp in [(1920, 1080, 4)] + ([(3840, 2160, 64)] if sys.argv[3] == "full" else []):
        best = 1e9
        for _ in range(2 if spp > 4 else 3):
            img, st = sc.render(rt.default_params(width=w, height=h, spp=spp))
            best = min(best, st.ms_total)
        out.append(f"{w}x{h}x{spp}: {best:.2f} ms {st.rays / best / 1e3:.1f} Mrays/s")
    print(" | ".join(out))
else:
    args = sys.argv[1:]
    full = "full" if args and args[0] == "full" else "probe"
    for v in [a for a in args if a != "full"]:
        name, _, envs = v.partition("@")
        env = dict(os.environ)
        for kv in filter(None, envs.split(",")):
            k, _, val = kv.partition("=")
            env[k] = val
        o = subprocess.run([sys.executable, __file__, "--child", str(ROOT / "gpurun_variants" / f"librt_{name}.so"), full], capture_output=True, text=True, env=env)
        print(v, o.stdout.strip() or o.stderr[-300:], flush=True)
