"""Multi-GPU host logic of the path: one process per GPU, rays shard naturally.

* training: every rank draws its own rays (nerfstudio DDP semantics, seed + rank) and the flat
  gradient arena is all-reduced (mean) once per step - the only exchange step of the path;
* render / eval: frames are dealt round-robin over ranks, no collective in the compute.

The functions take CPU or CUDA tensors and any backend, so the ``gloo`` world_size-2 tests on CPU
exercise exactly the code the NCCL runs use.
"""

from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor


def world_info(group=None) -> tuple:
    """(rank, world_size), (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_frames(num_frames: int, rank: int, world_size: int) -> List[int]:
    """Frame indices rendered by ``rank``: r, r + world, r + 2 world, ... (every frame exactly once)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, num_frames, world_size))


def rank_seed(base_seed: int, rank: int, step: int = 0) -> int:
    """Seed of the pixel batch ``step`` on ``rank`` (independent draws per rank, as nerfstudio's per-rank
    datamanagers)."""
    return int(base_seed) + 1000 * int(rank) + int(step)


def allreduce_mean_(flat: Tensor, group=None, world_size: Optional[int] = None) -> Tensor:
    """In-place mean over ranks of one flat buffer (DDP gradient semantics).  NCCL averages inside the
    collective; backends without ReduceOp.AVG (gloo) sum and scale."""
    if world_size is None:
        world_size = world_info(group)[1]
    if world_size <= 1:
        return flat
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world_size)
    return flat


def allreduce_mean_grads_(params: Sequence[Tensor], group=None, world_size: Optional[int] = None) -> None:
    """Mean over ranks of ``p.grad`` for the plugin (autograd) route.  Gradients that are views of one
    contiguous arena (what ``functional.render``'s backward produces) go out as a single collective."""
    if world_size is None:
        world_size = world_info(group)[1]
    if world_size <= 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    base = grads[0]._base if grads[0]._base is not None else None
    if base is not None and all(g._base is base for g in grads) and base.is_contiguous():
        span = sum((g.numel() + 3) // 4 * 4 for g in grads)
        if span == base.numel():
            allreduce_mean_(base, group, world_size)
            return
    for g in grads:
        allreduce_mean_(g, group, world_size)


def gather_frames(frames: Tensor, num_frames: int, group=None) -> Optional[Tensor]:
    """Rank 0 receives the frames rendered by every rank in frame order ([num_frames, ...]); other ranks get
    None.  ``frames`` holds this rank's share in the order of :func:`shard_frames`."""
    rank, world = world_info(group)
    if world == 1:
        return frames
    per = (num_frames + world - 1) // world
    pad = torch.zeros((per, *frames.shape[1:]), dtype=frames.dtype, device=frames.device)
    pad[: frames.shape[0]] = frames
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if rank != 0:
        return None
    out = torch.empty((num_frames, *frames.shape[1:]), dtype=frames.dtype, device=frames.device)
    for r in range(world):
        idx = shard_frames(num_frames, r, world)
        out[idx] = bufs[r][: len(idx)]
    return out
