"""Two real GPUs: sample-sharded render, then the fused peer-memory reduce + tonemap (NVLS multimem and plain
peer loads) against the NCCL reduce + single-GPU tonemap and against the one-GPU render of all samples."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["RT_ROOT"])
import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200.multi_gpu import sample_range
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
first, count = sample_range(SPP, rank, world)
dev = torch.device("cuda", rank)
with torch.cuda.stream(stream):
    accum = symm.empty((H, W, 4), dtype=torch.float32, device=dev)
    rgb = symm.empty((H, W, 3), dtype=torch.float32, device=dev)
    total = symm.empty((H, W, 4), dtype=torch.float32, device=dev)
    hdl, h_rgb, h_tot = (symm.rendezvous(t, dist.group.WORLD) for t in (accum, rgb, total))
    accum.zero_()
    scene.render_accum_device(rt.default_params(width=W, height=H, spp=count, sample_offset=first), accum.data_ptr())
    row0, row1 = rank * H // world, (rank + 1) * H // world
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
        ref_rgb = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
        rt.tonemap_device(ctx, ref.data_ptr(), W, H, ref_rgb.data_ptr(), 0)
        one = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
        scene.render_accum_device(rt.default_params(width=W, height=H, spp=SPP), one.data_ptr())
        stream.synchronize()
        out.update(nccl_sum=ref.cpu().numpy(), nccl_rgb=ref_rgb.cpu().numpy(), one_gpu_sum=one.cpu().numpy(),
                   has_multicast=np.array([int(mc_ptr != 0)]))
        np.savez(os.environ["RT_OUT"], **out)
dist.barrier()
dist.destroy_process_group()
"""


def test_two_gpu_fused_reduce_tonemap(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "out.npz"
    env = dict(os.environ, RT_ROOT=str(ROOT), RT_OUT=str(out))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                           "--master-port", str(port), str(script)], env=env, timeout=900)
    r = np.load(out)
    # two ranks: a + b is the same float sum whichever engine adds it
    assert np.array_equal(r["peer_sum"], r["nccl_sum"]) and np.array_equal(r["peer_rgb"], r["nccl_rgb"])
    if int(r["has_multicast"][0]):
        assert np.array_equal(r["multimem_sum"], r["nccl_sum"]) and np.array_equal(r["multimem_rgb"], r["nccl_rgb"])
    # and the sharded job traced exactly the paths of the one-GPU job (float atomics commute up to rounding)
    assert np.array_equal(r["nccl_sum"][..., 3], r["one_gpu_sum"][..., 3])
    assert np.allclose(r["nccl_sum"], r["one_gpu_sum"], rtol=1e-5, atol=1e-5)
