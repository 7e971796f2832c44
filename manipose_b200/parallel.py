"""Clip-sharded multi-GPU plumbing: one process per GPU, no collective on the data path.

Nothing in ``RMCLManifoldMixSTE.forward`` mixes clips (attention is within a frame or within one (clip, joint) track, bone
lengths are a per-clip mean, decoder and aggregation are per pose), so inference shards dim 0 across ranks and every rank
writes its own output shard.  The only exchange on the path is the scalar reduction of dataset-level metric sums
(hpe/eval_utils.py:165-185), one ``all_reduce(SUM)`` at the end.  ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests)
is plumbing here.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size) of the default process group, (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of ``n_items`` units for ``rank``: the first ``n_items % world_size`` ranks get one extra."""
    if rank is None or world_size is None:
        rank, world_size = world()
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_clips(x: torch.Tensor, rank: Optional[int] = None, world_size: Optional[int] = None) -> torch.Tensor:
    """This rank's clips (a view of dim 0)."""
    start, stop = shard_range(x.shape[0], rank, world_size)
    return x[start:stop]


def reduce_metric_sums(sums: torch.Tensor) -> torch.Tensor:
    """Sum partial metric accumulators (e.g. [sum of per-joint errors, number of frames]) over ranks, in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


def gather_clips(local: torch.Tensor, n_total: int) -> Optional[torch.Tensor]:
    """Optional host-side concat of per-rank output shards on rank 0 (shards may differ by one clip)."""
    rank, world_size = world()
    if world_size == 1:
        return local
    sizes = [shard_range(n_total, r, world_size) for r in range(world_size)]
    pad = max(b - a for a, b in sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world_size)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    return torch.cat([o[:b - a] for o, (a, b) in zip(out, sizes)], dim=0)
