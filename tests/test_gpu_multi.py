"""Real GPUs (gpurun --gpus N): the multi-GPU entries of the C-ABI.
  * rt_group_* — one process per GPU (torchrun), CUDA IPC peer mappings, flag barriers, fused peer-load reduce + tonemap:
    at 2, 4 and 8 ranks the root's image must be the finalisation of the rank-ordered float sum of the members'
    accumulators, bit for bit, and the sharded job must have traced exactly the paths of the one-GPU job;
    the NCCL reduce and (where the fabric has it) the NVLS multimem form are the A/B.
  * rt_multi_* — one process driving every visible device (the form the C++ application uses)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import raytracing_renderer_cuda_b200 as rt
from tests.conftest import ROOT, record_parity

pytestmark = pytest.mark.gpu

WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["RT_ROOT"])
import raytracing_renderer_cuda_b200 as rt

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
W, H, SPP = 320, 180, 36
desc = rt.SceneDesc.builtin("book1_final")
ctx = rt.Context(rank)
scene = rt.Scene(ctx, desc)

def exchange(blob):
    table = [None] * world
    dist.all_gather_object(table, blob)
    return table

g = rt.Group(ctx, rank, world, W, H, exchange)
first, count = rt.shard_samples(SPP, rank, world)
frames = []
for frame in range(2):                       # twice: the second frame re-uses accumulators the peers have read
    g.begin_frame()
    scene.render_accum_device(rt.default_params(width=W, height=H, spp=count, sample_offset=first), g.accum_ptr)
    ms = g.finish_frame(want_rgb8=True, timed=(frame == 1))
    rgb, rgb8 = np.empty((H, W, 3), np.float32), np.empty((H, W, 3), np.uint8)
    g.read_frame(rgb if rank == 0 else None, rgb8 if rank == 0 else None)
    acc = g.read_accum()
    accs = [None] * world
    dist.all_gather_object(accs, acc)
    frames.append((rgb, rgb8, accs, ms))
out = {}
if rank == 0:
    for k, (rgb, rgb8, accs, ms) in enumerate(frames):
        out[f"rgb{k}"], out[f"rgb8_{k}"], out[f"accs{k}"], out[f"ms{k}"] = rgb, rgb8, np.stack(accs), np.array([ms])
    # A/B 1: NCCL reduce of the same accumulators + the single-GPU finalisation
    dev = torch.device("cuda", rank)
t = torch.from_numpy(frames[1][2][rank]).cuda()
dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
if rank == 0:
    ref_rgb = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    rt.tonemap_device(ctx, t.data_ptr(), W, H, ref_rgb.data_ptr(), 0)
    ctx.synchronize()
    out["nccl_sum"], out["nccl_rgb"] = t.cpu().numpy(), ref_rgb.cpu().numpy()
    one, st = scene.render_accum(rt.default_params(width=W, height=H, spp=SPP))
    out["one_gpu_sum"] = one
    np.savez(os.environ["RT_OUT"], **out)
g.close()
dist.barrier()
dist.destroy_process_group()
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 4, 8])
def test_group_fused_reduce_tonemap(tmp_path, oracle, world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "out.npz"
    env = dict(os.environ, RT_ROOT=str(ROOT), RT_OUT=str(out))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                           "--master-port", str(_free_port()), str(script)], env=env, timeout=900)
    r = np.load(out)
    for k in (0, 1):
        accs = r[f"accs{k}"]
        total = accs[0].copy()
        for a in accs[1:]:
            total = total + a  # float32, rank order: what the kernel does
        want = oracle.tonemap(total)
        assert np.array_equal(r[f"rgb{k}"], want), k                       # bit for bit
        assert np.array_equal(r[f"rgb8_{k}"], rt.quantize_rgb8(want)), k
    # the sharded job traced exactly the paths of the one-GPU job (float atomics commute up to rounding)
    total = r["accs1"][0].copy()
    for a in r["accs1"][1:]:
        total = total + a
    assert np.array_equal(total[..., 3], r["one_gpu_sum"][..., 3])
    assert np.allclose(total, r["one_gpu_sum"], rtol=1e-5, atol=1e-5)
    # NCCL's reduction tree may add in another order: equal at 2 ranks, equal up to rounding beyond
    if world == 2:
        assert np.array_equal(r["nccl_sum"], total) and np.array_equal(r["nccl_rgb"], r["rgb1"])
    else:
        assert np.allclose(r["nccl_sum"], total, rtol=1e-6, atol=1e-6)
    record_parity("multi_gpu_group", world=world, image_bit_exact=1, ms_barrier_reduce_barrier=float(r["ms1"][0]),
                  max_abs_vs_one_gpu=float(np.abs(total - r["one_gpu_sum"]).max()))


def test_single_process_multi_device(oracle):
    """rt_multi_*: every visible device from ONE process (the C++ application's --gpus path)."""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    W, H, SPP = 320, 180, 36
    desc = rt.SceneDesc.builtin("book1_final")
    m = rt.Multi()
    assert m.size == n
    m.set_scene(desc)
    p = rt.default_params(width=W, height=H, spp=SPP)
    for _ in range(2):
        rgb, rgb8, st, ms = m.render(p, want_rgb8=True)
        accs = [m.read_accum(k, W, H) for k in range(n)]
        total = accs[0].copy()
        for a in accs[1:]:
            total = total + a
        want = oracle.tonemap(total)
        assert np.array_equal(rgb, want) and np.array_equal(rgb8, rt.quantize_rgb8(want))
        assert np.array_equal(total[..., 3], np.full((H, W), SPP, np.float32))
        assert st.paths == W * H * SPP and st.rays > st.paths
    one, st1 = rt.Scene(rt.Context(0), desc).render_accum(p)
    assert st.rays == st1.rays
    assert np.allclose(total, one, rtol=1e-5, atol=1e-5)
    record_parity("multi_gpu_single_process", devices=n, image_bit_exact=1, ms_reduce=ms)
    m.close()


SYMM_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["RT_ROOT"])
import raytracing_renderer_cuda_b200 as rt
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
W, H, SPP = 320, 180, 32
desc = rt.SceneDesc.builtin("book1_final")
ctx = rt.Context(rank)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
scene = rt.Scene(ctx, desc)
first, count = rt.shard_samples(SPP, rank, world)
dev = torch.device("cuda", rank)
with torch.cuda.stream(stream):
    accum = symm.empty((H, W, 4), dtype=torch.float32, device=dev)
    rgb = symm.empty((H, W, 3), dtype=torch.float32, device=dev)
    total = symm.empty((H, W, 4), dtype=torch.float32, device=dev)
    hdl, h_rgb, h_tot = (symm.rendezvous(t, dist.group.WORLD) for t in (accum, rgb, total))
    accum.zero_()
    scene.render_accum_device(rt.default_params(width=W, height=H, spp=count, sample_offset=first), accum.data_ptr())
    row0, row1 = rt.shard_rows(H, rank, world)
    out = {}
    mc_ptr = int(hdl.multicast_ptr or 0)  # 0 when the fabric has no NVLS multicast
    for name, mc in (("peer", 0), ("multimem", mc_ptr)):
        rgb.zero_(); total.zero_()
        hdl.barrier(0)
        rt.reduce_tonemap_peers(ctx, [int(p) for p in hdl.buffer_ptrs], mc, W, H, row0, row1,
                                int(h_rgb.buffer_ptrs[0]), 0, int(h_tot.buffer_ptrs[0]))
        hdl.barrier(1)
        stream.synchronize()
        if rank == 0:
            out[name + "_rgb"] = rgb.cpu().numpy()
            out[name + "_sum"] = total.cpu().numpy()
    ref = accum.clone()
    dist.reduce(ref, dst=0, op=dist.ReduceOp.SUM)
    stream.synchronize()
    if rank == 0:
        out.update(nccl_sum=ref.cpu().numpy(), has_multicast=np.array([int(mc_ptr != 0)]))
        np.savez(os.environ["RT_OUT"], **out)
dist.barrier()
dist.destroy_process_group()
"""


def test_two_gpu_nvls_multimem_variant(tmp_path):
    """The A/B form of the same kernel on torch's symmetric memory: NVLS `multimem.ld_reduce` (the switch adds the copies)
    against peer loads added in rank order, both against NCCL."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(SYMM_WORKER)
    out = tmp_path / "out.npz"
    env = dict(os.environ, RT_ROOT=str(ROOT), RT_OUT=str(out))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                           "--master-port", str(_free_port()), str(script)], env=env, timeout=900)
    r = np.load(out)
    assert np.array_equal(r["peer_sum"], r["nccl_sum"])  # two ranks: a + b is the same float sum whichever engine adds it
    if int(r["has_multicast"][0]):
        assert np.array_equal(r["multimem_sum"], r["nccl_sum"]) and np.array_equal(r["multimem_rgb"], r["peer_rgb"])
