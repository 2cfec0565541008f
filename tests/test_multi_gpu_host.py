"""Host logic of the sample-sharded multi-GPU path (rt_shard_samples / rt_shard_rows of the C-ABI), exercised on CPU
with two gloo ranks: each rank renders its sample range (the oracle in product-sampler mode stands in for the GPU), the
float4 accumulators are summed onto rank 0, and the result must equal the single-rank render."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import raytracing_renderer_cuda_b200 as rt
from raytracing_renderer_cuda_b200 import capi
from tests.conftest import ROOT


def test_sample_ranges_partition_the_samples():
    for total in (0, 1, 7, 100, 4096):
        for world in (1, 2, 3, 8, 16):
            got = [rt.shard_samples(total, r, world) for r in range(world)]
            assert sum(c for _, c in got) == total
            nxt = 0
            for first, count in got:
                assert first == nxt and count >= 0
                nxt += count
            assert max(c for _, c in got) - min(c for _, c in got) <= 1
    with pytest.raises(capi.RtError):
        rt.shard_samples(10, 2, 2)
    with pytest.raises(capi.RtError):
        rt.shard_samples(-1, 0, 2)


def test_row_bands_partition_the_frame():
    for height in (0, 1, 7, 600, 4320):
        for world in (1, 2, 3, 8, 16):
            bands = [rt.shard_rows(height, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == height
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0 and a0 <= a1 and b0 <= b1
    with pytest.raises(capi.RtError):
        rt.shard_rows(10, 3, 3)


def test_multi_gpu_entries_need_a_device():
    """No CPU fallback: without a CUDA device the single-process entry reports RT_ERR_NO_DEVICE."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(capi.RtError) as e:
        rt.Multi()
    assert e.value.status == capi.RT_ERR_NO_DEVICE


WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["RT_ROOT"])
import raytracing_renderer_cuda_b200 as rt
from tests.oracle_api import Oracle

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
desc = rt.SceneDesc.builtin("book1_final")
sc = Oracle().scene(desc)
W, H, SPP = 40, 24, 6
first, count = rt.shard_samples(SPP, rank, world)
p = rt.default_params(width=W, height=H, spp=count, sample_offset=first)
acc, _ = sc.render(p, sampler=1, arith=1, nthreads=2)
acc = torch.from_numpy(acc)
dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)
if rank == 0:
    np.save(os.environ["RT_OUT"], acc.numpy())
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_gloo_reduce_equals_single_rank(tmp_path, oracle):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "acc.npy"
    env = dict(os.environ, RT_ROOT=str(ROOT), RT_OUT=str(out), OMP_NUM_THREADS="1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                           "--master-port", str(port), str(script)], env=env, timeout=600)
    got = np.load(out)
    sc = oracle.scene(rt.SceneDesc.builtin("book1_final"))
    want, _ = sc.render(rt.default_params(width=40, height=24, spp=6), sampler=1, arith=1, nthreads=2)
    assert np.array_equal(got[..., 3], want[..., 3])                # 6 samples per pixel in total
    assert np.allclose(got, want, rtol=1e-6, atol=1e-6)              # same paths; only the addition order differs
