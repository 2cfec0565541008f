"""Sample-sharded multi-GPU rendering: one process per GPU (torch.distributed), each rank renders a disjoint
range of sample indices of EVERY pixel into its own float4 accumulator, one reduce(SUM) onto rank 0, tonemap.

The Philox keys use the global sample index (rt_render_params.sample_offset), so the set of paths traced — and
therefore the image, up to float addition order — does not depend on the number of GPUs.  The path has no other
exchange step, so there is no other collective."""
from __future__ import annotations

from typing import Callable, Tuple


def sample_range(total_spp: int, rank: int, world: int) -> Tuple[int, int]:
    """(first sample index, sample count) of `rank`: contiguous ranges, the remainder spread over the low ranks."""
    if world < 1 or not 0 <= rank < world or total_spp < 0:
        raise ValueError("bad rank/world/total_spp")
    base, rem = divmod(total_spp, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def render_sharded(render_range: Callable[[int, int], "object"], total_spp: int, rank: int, world: int, reduce_to_root=None):
    """Renders this rank's share.  `render_range(first, count)` returns the rank's accumulator (any tensor type the
    caller's `reduce_to_root` understands); `reduce_to_root(acc)` sums the accumulators onto rank 0 in place
    (torch.distributed.reduce with the NCCL backend on GPUs, gloo in the CPU tests)."""
    first, count = sample_range(total_spp, rank, world)
    acc = render_range(first, count)
    if world > 1 and reduce_to_root is not None:
        reduce_to_root(acc)
    return acc


def nccl_reduce_to_root(acc) -> None:
    import torch.distributed as dist

    dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)
