#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the per-pixel render path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c1|c2|c3|c4|c5] [--impl ours|reference]

A "step" is one frame of the configured scene: zero the float4 accumulator, render `spp`
samples per pixel into it (wavefront pipeline, librt_b200.so), [N > 1: sum the accumulators of
all ranks onto rank 0 with one NCCL reduce over NVLink], tonemap (/spp, saturate, sqrt) and
quantise on the device.  N > 1 shards SAMPLES: rank r renders sample indices [r*spp, (r+1)*spp)
of every pixel, so the job renders N*spp samples per pixel ("weak" scaling: per-GPU work fixed).

`value` is device-timed (CUDA events on the stream the kernels run on) with the scene resident
in HBM; `e2e` goes through the C-ABI with HOST buffers: scene upload (H2D) + render + read-back
of the finished frame (D2H) inside the timed region.  `--impl reference` times the reference's
own main.cu rebuilt unchanged for sm_100 (oracle/_ref/ref_main: BASELINE.json's stated baseline —
the reference ships no CPU renderer) for c1, and the reference harness for the other configs.
The `cpu_baseline` leg times the reference headers compiled for the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# config -> (builtin scene, width, height, spp per step and GPU, sphere count for random_spheres)
CONFIGS = {
    "c1": dict(scene="earth_emitter", width=1200, height=600, spp=100, n=0,
               name="C1 earth_emitter 1200x600x100spp depth 50 (reference main.cu:188-356)"),
    "c2": dict(scene="book1_final", width=1920, height=1080, spp=256, n=0, name="C2 book-1 final ~485 spheres 1920x1080x256spp"),
    "c3": dict(scene="perlin_motion", width=1200, height=600, spp=1024, n=0,
               name="C3 perlin/checker/wood + moving spheres + emitters 1200x600x1024spp"),
    "c4": dict(scene="random_spheres", width=3840, height=2160, spp=64, n=1_000_000, name="C4 1M random spheres (GPU LBVH) 3840x2160x64spp"),
    "c5": dict(scene="book1_final", width=7680, height=4320, spp=16, n=0,
               name="C5 book-1 final 7680x4320, 16 spp per step and GPU (4096 spp = 256 steps)"),
}
FLOP_PER_RAY = {"c1": 420.0, "c2": 420.0, "c3": 615.0, "c4": 815.0, "c5": 420.0}  # SURVEY.md §8(d), frozen
STATE_BYTES_PER_RAY = 120.0


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": float(d["hbm_gbs"]), "sm_max_mhz": float(d.get("sm_max_mhz", 1965.0)), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(cfg_key: str, desc, target_s: float = 12.0) -> dict:
    """Host-core baseline on a bounded sample of the same workload: the reference's own headers compiled
    for the host (oracle/_ref/libref_cpu_rn.so, kind "reference"), else the oracle port (kind "port")."""
    from tests import oracle_api as oa

    cfg = CONFIGS[cfg_key]
    cores = os.cpu_count() or 1
    w, h = cfg["width"] // 4, cfg["height"] // 4

    import raytracing_renderer_cuda_b200 as rt

    if oa.REFCPU_RN_SO.exists():
        kind = "reference"
        sc = oa.RefCpu(oa.REFCPU_RN_SO).scene(desc, use_bvh=True)
        run = lambda spp: sc.render(w, h, spp, nthreads=cores, want_fb=False)[2]  # noqa: E731
    else:
        kind = "port"
        sc = oa.Oracle().scene(desc)
        run = lambda spp: sc.render(rt.default_params(width=w, height=h, spp=spp), sampler=0, arith=0, nthreads=cores)[1]  # noqa: E731
    t0 = time.perf_counter()
    run(1)
    probe = max(time.perf_counter() - t0, 1e-4)  # one sample per pixel at 1/16 of the frame
    rate = w * h / probe
    if rate * target_s > 4 * w * h * 4:  # fast enough: use the full frame
        w, h = cfg["width"], cfg["height"]
        sc = None
        if kind == "reference":
            sc = oa.RefCpu(oa.REFCPU_RN_SO).scene(desc, use_bvh=True)
            run = lambda spp: sc.render(w, h, spp, nthreads=cores, want_fb=False)[2]  # noqa: E731
        else:
            sc = oa.Oracle().scene(desc)
            run = lambda spp: sc.render(rt.default_params(width=w, height=h, spp=spp), sampler=0, arith=0, nthreads=cores)[1]  # noqa: E731
    spp = int(max(1, min(cfg["spp"], rate * target_s / (w * h))))
    t0 = time.perf_counter()
    rays = run(spp)
    dt = time.perf_counter() - t0
    paths = w * h * spp
    return {"value": paths / dt / 1e6, "unit": "Mpaths/s", "cores": cores, "kind": kind, "mrays_per_s": rays / dt / 1e6,
            "seconds": dt, "sample": f"{cfg['scene']} {w}x{h}x{spp}spp ({paths} paths) on {cores} host threads; "
            + ("reference headers via oracle/shim, _rz intrinsics rounding to nearest" if kind == "reference" else "oracle/rt_oracle.cpp")}


def build_desc(cfg_key: str):
    import raytracing_renderer_cuda_b200 as rt
    from raytracing_renderer_cuda_b200.assets import load_earth

    cfg = CONFIGS[cfg_key]
    image = load_earth() if cfg["scene"] == "earth_emitter" else None
    return rt.SceneDesc.builtin(cfg["scene"], image, n=cfg["n"])


def desc_h2d_bytes(desc) -> int:
    from raytracing_renderer_cuda_b200 import capi

    d = desc.desc
    b = d.n_spheres * C.sizeof(capi.rt_sphere) + d.n_materials * C.sizeof(capi.rt_material) + d.n_textures * C.sizeof(capi.rt_texture)
    for i in range(d.n_images):
        b += d.images[i].width * d.images[i].height * 3 * 4
    return int(b)


# ------------------------------------------------------------------------------------------ ours
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import raytracing_renderer_cuda_b200 as rt
    from raytracing_renderer_cuda_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N > 1 through torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; librt_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = CONFIGS[args.config]
    W, H = cfg["width"], cfg["height"]
    spp = args.spp or cfg["spp"]
    desc = build_desc(args.config)
    ctx = rt.Context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    scene = rt.Scene(ctx, desc)
    info = scene.info()
    params = rt.default_params(width=W, height=H, spp=spp, sample_offset=rank * spp)

    with torch.cuda.stream(stream):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
        collective = "none"
        hdl = None
        if world > 1 and args.collective != "nccl":
            # symmetric memory: every rank maps every rank's accumulator (NVLink peer pointers + NVLS multicast)
            try:
                import torch.distributed._symmetric_memory as symm

                accum = symm.empty((H, W, 4), dtype=torch.float32, device=torch.device("cuda", local_rank))
                rgb = symm.empty((H, W, 3), dtype=torch.float32, device=torch.device("cuda", local_rank))
                rgb8 = symm.empty((H, W, 3), dtype=torch.uint8, device=torch.device("cuda", local_rank))
                hdl = symm.rendezvous(accum, dist.group.WORLD)
                h_rgb = symm.rendezvous(rgb, dist.group.WORLD)
                h_rgb8 = symm.rendezvous(rgb8, dist.group.WORLD)
                peer_ptrs = [int(p) for p in hdl.buffer_ptrs]
                mc_ptr = int(hdl.multicast_ptr or 0) if args.collective in ("auto", "multimem") else 0  # 0: no NVLS multicast
                root_rgb, root_rgb8 = int(h_rgb.buffer_ptrs[0]), int(h_rgb8.buffer_ptrs[0])
                row0, row1 = rank * H // world, (rank + 1) * H // world
                collective = "fused peer-memory reduce+tonemap (" + ("NVLS multimem.ld_reduce" if mc_ptr else "NVLink peer loads") + ")"
            except Exception as ex:  # noqa: BLE001
                if args.collective != "auto":
                    raise
                print(f"[bench] symmetric memory unavailable ({ex!r}); using the NCCL reduce", file=sys.stderr)
                hdl = None
        if hdl is None:
            accum = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
            rgb = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
            rgb8 = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
            if world > 1:
                collective = "NCCL reduce(SUM) to rank 0 + tonemap"

        def finish():
            """sum over ranks + pixel finalisation; the frame ends up on rank 0"""
            if hdl is not None:
                hdl.barrier(0)  # every rank has finished rendering into its accumulator
                rt.reduce_tonemap_peers(ctx, peer_ptrs, mc_ptr, W, H, row0, row1, root_rgb, root_rgb8)
                hdl.barrier(1)  # every band is in rank 0's image; accumulators may be reused
            else:
                if world > 1:
                    dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
                if rank == 0:
                    rt.tonemap_device(ctx, accum.data_ptr(), W, H, rgb.data_ptr(), rgb8.data_ptr())

        def step():
            accum.zero_()
            scene.render_accum_device(params, accum.data_ptr())
            finish()

        # one instrumented step: rays and launches per step (deterministic: the RNG is keyed on pixel/sample/bounce)
        accum.zero_()
        st = scene.render_accum_device(params, accum.data_ptr(), want_stats=True)
        rays_rank, launches_step, iters = int(st.rays), int(st.launches) + 1, int(st.iterations)
        for _ in range(args.warmup):
            step()
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        rv = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        t_wall0 = time.perf_counter()
        for k in range(args.steps):
            flush.fill_(k & 255)  # L2 flush between timed iterations, outside the timed span
            ev[k][0].record(stream)
            accum.zero_()
            rv[k][0].record(stream)
            scene.render_accum_device(params, accum.data_ptr())
            rv[k][1].record(stream)
            finish()
            ev[k][1].record(stream)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.stop() if sampler else None
        ms_steps = [a.elapsed_time(b) for a, b in ev]
        ms_render = [a.elapsed_time(b) for a, b in rv]
        ms_step = float(np.mean(ms_steps))
        ms_kernel_span = float(np.mean(ms_render))

        # ---- end to end through the C-ABI with HOST buffers: scene upload + render + read-back each step ----
        h2d = desc_h2d_bytes(desc)
        out_host = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
        acc_host = torch.empty((H, W, 4), dtype=torch.float32).pin_memory() if world > 1 else None
        cudart = torch.cuda.cudart()
        for i in range(desc.desc.n_images):  # pin the host image the scene upload reads
            im = desc.desc.images[i]
            cudart.cudaHostRegister(C.cast(im.rgb, C.c_void_p).value, im.width * im.height * 12, 0)

        def e2e_step():
            sc = rt.Scene(ctx, desc)  # H2D: spheres, materials, textures, image; BVH build
            if world == 1:
                sc.render(params, out_host.numpy())  # render + tonemap + D2H, synchronous
            else:
                accum.zero_()
                sc.render_accum_device(params, accum.data_ptr())
                finish()
                if rank == 0:
                    out_host.copy_(rgb, non_blocking=True)
                stream.synchronize()
            sc.close()

        for _ in range(min(args.warmup, 2)):
            e2e_step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.steps

        # ---- output stage on the device: render + finalise + flip/quantise + JPEG, only the file is read back ----
        jpeg = None
        if world == 1:
            sc = rt.Scene(ctx, desc)
            file_host = torch.empty(int(ctx.lib.rt_jpeg_max_bytes(W, H)), dtype=torch.uint8).pin_memory()
            for _ in range(min(args.warmup, 2)):
                sc.render_jpeg(params, 100, file_host.numpy())
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                f_bytes, st_j = sc.render_jpeg(params, 100, file_host.numpy())
            t_j = (time.perf_counter() - t0) / args.steps
            sc.close()
            jpeg = {"value": float(W) * H * spp / t_j / 1e6, "unit": "Mpaths/s", "ms_per_step": t_j * 1e3,
                    "ms_device_jpeg": float(st_j.ms_d2h), "d2h_bytes_per_step": int(f_bytes.size), "quality": 100,
                    "gb_per_s_pixels": W * H * 3 / (float(st_j.ms_d2h) / 1e3) / 1e9 if st_j.ms_d2h > 0 else None,
                    "what": "rt_render_jpeg: render + finalise + Y-flip/quantise + baseline JPEG (byte-identical to the reference's "
                            "stbi_write_jpg, main.cu:475-491) on the device; the D2H copy moves the finished file only"}

    vals = torch.tensor([ms_step, ms_kernel_span, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    rays_t = torch.tensor([rays_rank], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays_t, op=dist.ReduceOp.SUM)
    ms_step, ms_kernel_span, e2e_ms = (float(x) for x in vals.tolist())
    rays_total = float(rays_t.item())

    if rank == 0:
        peaks = load_peaks()
        paths_total = float(W) * H * spp * world
        value = paths_total / ms_step / 1e3
        n_step_launches = max(1, launches_step - 2)  # k_wf_step launches (excl. k_wf_init, tonemap)
        kernel_name = ("k_wf_step_pt" if int(info.n_spheres) >= 4096 else "k_wf_step_warp") if int(info.n_nodes) else "k_wf_step_cta"
        grain = os.environ.get("RT_WF_GRAIN", "")
        if grain.startswith("r") and iters == 1:  # experimental single-launch kernel: k_ring_fill + k_ring_commit + k_wf_ring
            kernel_name, n_step_launches = "k_wf_ring", 1
        fp32_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12  # TFLOP/s at the measured max SM clock
        flop_launch = rays_rank / n_step_launches * FLOP_PER_RAY[args.config]
        dur_launch_s = ms_kernel_span / 1e3 / n_step_launches
        achieved = flop_launch / dur_launch_s / 1e12
        hbm_achieved = rays_rank * STATE_BYTES_PER_RAY / (ms_kernel_span / 1e3) / 1e9
        traffic = None  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
        ncu_static = None  # what that capture says about issue utilisation / lanes / stalls (static: not measured by this run)
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            ent = json.loads(tp.read_text()).get(args.config) or {}
            traffic = ent.get("bytes_per_launch")
            ncu_static = ent.get("ncu")
        line = {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "mrays_per_s": rays_total / ms_step / 1e3, "ms_per_frame": ms_step,
            "rays_per_path": rays_total / paths_total,
            "config": {"workload": cfg["name"], "scene": cfg["scene"], "width": W, "height": H, "spp_per_gpu": spp,
                       "spp_total": spp * world, "max_depth": 50, "n_spheres": int(info.n_spheres), "bvh_nodes": int(info.n_nodes),
                       "bvh_mode": int(info.bvh_mode), "pipeline": "wavefront" + (f" (RT_WF_GRAIN={os.environ['RT_WF_GRAIN']})" if os.environ.get("RT_WF_GRAIN") else ""), "parallelism": f"samples x{world}", "collective": collective,
                       "l2": "flushed between timed steps (256 MiB fill, outside the timed spans)",
                       "timing": "sum of per-step CUDA-event spans on the launching stream, max over ranks"},
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                         "traffic": traffic, "ncu": ncu_static, "kernel": kernel_name, "launches_per_step": n_step_launches,
                         "avg_launch_ms": dur_launch_s * 1e3, "flop_per_ray": FLOP_PER_RAY[args.config],
                         "peak_source": f"148 SM x 128 lanes x 2 flop x {peaks['sm_max_mhz']:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz, {peaks['source']})",
                         "note": "no dense contraction and L2-resident state: the bounding roofline is FP32 issue (SURVEY.md 8d), not hbm/tensor",
                         "hbm": {"achieved": hbm_achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_achieved / peaks["hbm_gbs"],
                                 "bytes_per_ray": STATE_BYTES_PER_RAY, "peak_source": f"MEASURED_PEAKS.json ({peaks['source']})"}},
            "e2e": {"value": paths_total / e2e_ms / 1e3, "unit": "Mpaths/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": W * H * 3 * 4,
                    "what": "rt_scene_create (H2D scene + texture, BVH build) + rt_render to a pinned host buffer, wall clock"},
            "gpu_launches": int(args.steps * launches_step),  # k_wf_init + k_wf_step x iterations + tonemap / fused reduce-tonemap
            "wavefront_iterations": iters, "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        if jpeg is not None:
            line["output_stage"] = jpeg
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.config, desc)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ reference
def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the reference is single-GPU: rank 0 alone runs and prints it
    cfg = CONFIGS[args.config]
    ref = ROOT / "oracle" / "_ref"
    W, H = cfg["width"], cfg["height"]
    desc = build_desc(args.config)
    cb = None if args.no_cpu_baseline else cpu_baseline(args.config, desc)
    if args.ref_device == "cpu":
        v = cb or cpu_baseline(args.config, desc)
        line = {"impl": "reference", "metric": "Mpaths/s", "value": v["value"], "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": 1,
                "warmup": 0, "ms_per_step": v["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": {"workload": cfg["name"], "device": "host cores"},
                "cpu_baseline": v, "e2e": {"value": v["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    if not (ref / "ref_main").exists() or not (ref / "ref_harness").exists():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_main not built (needs /root/reference at build time)"}))
        return
    spp = args.spp or cfg["spp"]
    took_ms, kernel_ms, rays = [], None, None
    if args.config == "c1" and not args.spp:
        # the UNCHANGED binary: its own timing window (init_rand_state + render + syncs, main.cu:431-454)
        for k in range(args.warmup + args.steps):
            o = subprocess.run([str(ref / "ref_main")], cwd=str(ref), capture_output=True, text=True, timeout=600)
            m = re.search(r"took (\d+)us", o.stdout)
            if o.returncode != 0 or not m:
                print(json.dumps({"impl": "reference", "unavailable": f"ref_main failed rc={o.returncode}: {o.stderr[-200:]}"}))
                return
            if k >= args.warmup:
                took_ms.append(int(m.group(1)) / 1e3)
        how = "reference src/main.cu rebuilt unchanged (-arch=sm_100), its own chrono window (init_rand_state + render + syncs)"
    with tempfile.TemporaryDirectory() as td:
        sp = Path(td) / "scene.rtsc"
        desc.save(str(sp))
        if args.config == "c4":
            cap = 10_000  # the reference builds its BVH on ONE device thread, O(N log^2 N) with virtual calls (bvh.h:75-113):
            #               10^5 spheres did not finish in 20 minutes on a B200, 10^6 is out of reach
            import raytracing_renderer_cuda_b200 as rt

            rt.SceneDesc.builtin("random_spheres", n=cap).save(str(sp))
        o = subprocess.run([str(ref / "ref_harness"), "render", str(sp), str(W), str(H), str(spp), str(Path(td) / "o.f32"), "1", "1",
                            str(max(1, min(args.steps, 3)))], capture_output=True, text=True, timeout=900)
        if o.returncode == 0:
            info = json.loads(o.stdout.strip().splitlines()[-1])
            kernel_ms, rays = info["ms_render"] + info["ms_init_rand"], info["rays"]
    if not took_ms:
        if kernel_ms is None:
            print(json.dumps({"impl": "reference", "unavailable": "ref_harness failed: " + o.stderr[-200:]}))
            return
        took_ms = [kernel_ms]
        how = "reference translation unit through oracle/ref_harness.cu (runtime sizes), CUDA events around init_rand_state + render"
    ms = float(np.mean(took_ms))
    paths = float(W) * H * spp
    value = paths / ms / 1e3
    line = {"impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "ms_per_frame": ms, "mrays_per_s": (rays / ms / 1e3) if rays else None,
            "kernel_ms_cuda_events": kernel_ms,
            "config": {"workload": cfg["name"], "device": "1x B200 (the reference is single-GPU; it ships no CPU renderer)", "how": how,
                       "spp": spp},
            "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if cb:
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)


def main() -> None:
    # stdout carries exactly ONE JSON line: anything libraries print on fd 1 (NCCL's version banner, ...) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", choices=sorted(CONFIGS), default="c1")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel per step and GPU")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--ref-device", choices=["gpu", "cpu"], default="gpu")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", choices=["auto", "multimem", "peer", "nccl"], default="auto",
                    help="N > 1: fused peer-memory reduce+tonemap (NVLS multimem / peer loads) or the plain NCCL reduce")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), str(Path(__file__).resolve())] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
