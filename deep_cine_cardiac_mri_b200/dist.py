"""Data-parallel sharding of cine slices over the GPUs of one box.

The SENSE path has no exchange step in inference: volumes are independent, so
each rank owns volumes `i = rank, rank + world, ...` (what the reference's
VolumeSampler does for validation, data/volume_sampler.py:63-90) and results
are gathered only for reporting.  NCCL is used on GPUs, gloo in CPU tests.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the process group if world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_indices(n_volumes: int, rank: int, world: int) -> List[int]:
    """volume index i -> rank i mod world."""
    return list(range(rank, n_volumes, world))


def shard_counts(n_volumes: int, world: int) -> List[int]:
    return [len(range(r, n_volumes, world)) for r in range(world)]


def gather_volumes(local: Sequence[torch.Tensor], n_volumes: int, rank: int, world: int):
    """Reassemble per-volume results in volume order on every rank (reporting only)."""
    if world == 1:
        return list(local)
    objs = [None] * world
    dist.all_gather_object(objs, [t.cpu() for t in local])
    out = [None] * n_volumes
    for r, lst in enumerate(objs):
        for j, i in enumerate(shard_indices(n_volumes, r, world)):
            out[i] = lst[j]
    return out


def max_over_ranks(value: float, device=None) -> float:
    """Timing reduction of the bench contract: max over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
