"""A/B timing of library variants (tools/build_variant.sh) on the GPU box.
    python tools/ab_test.py base mb3 ...         # each variant in its own process
Env RT_WF_POOL is passed through."""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, str(ROOT))
    from raytracing_renderer_cuda_b200 import capi
    capi.LIB_PATH = Path(sys.argv[2])
    import raytracing_renderer_cuda_b200 as rt
    from raytracing_renderer_cuda_b200.assets import load_earth
    ctx = rt.Context(0)
    out = {}
    cases = [("c1", "earth_emitter", dict(image=load_earth()), (1200, 600, 100)), ("c2", "book1_final", {}, (960, 540, 64)),
             ("c3", "perlin_motion", {}, (600, 300, 64))]
    only = os.environ.get("AB_CASES", "c1,c2,c3").split(",")
    for key, name, kw, (w, h, spp) in cases:
        if key not in only: continue
        sc = rt.Scene(ctx, rt.SceneDesc.builtin(name, **kw))
        for pipe, pn in ((capi.RT_PIPE_WAVEFRONT, "wf"), (capi.RT_PIPE_MEGAKERNEL, "mega")):
            if pn == "mega" and os.environ.get("AB_NO_MEGA"): continue
            best = 1e9
            for _ in range(4):
                _, st = sc.render(rt.default_params(width=w, height=h, spp=spp, pipeline=pipe))
                best = min(best, st.ms_total)
            out[f"{key}_{pn}"] = round(best, 3)
            out[f"{key}_{pn}_it"] = st.iterations
    print(json.dumps(out))
else:
    for v in sys.argv[1:]:  # name[@ENV=value[,ENV=value...]]
        name, _, envs = v.partition("@")
        lib = ROOT / "gpurun_variants" / f"librt_{name}.so"
        env = dict(os.environ)
        for kv in filter(None, envs.split(",")):
            k, _, val = kv.partition("=")
            env[k] = val
        o = subprocess.run([sys.executable, __file__, "--child", str(lib)], capture_output=True, text=True, env=env)
        print(v, os.environ.get("RT_WF_POOL", ""), o.stdout.strip() or o.stderr[-400:], flush=True)
