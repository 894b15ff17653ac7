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


def bind_host_memory_to_gpu(local_rank: int) -> dict:
    """Best-effort NUMA placement for a rank's pinned host buffers: prefer the memory node (and the CPUs) the GPU hangs
    off, so that the sparse k-space upload (`ops.upload_masked_kspace`, the GPU reads pinned memory in place) and the
    result download do not cross the inter-socket link when 8 ranks feed 8 GPUs from one host.  Call before allocating
    pinned memory.  Returns what was done ({"node", "cpus", "mempolicy"}); never raises - on a box without NUMA
    information, or inside a container whose cpuset forbids it, nothing changes."""
    import ctypes
    import platform
    from pathlib import Path
    info = {"node": None, "cpus": None, "mempolicy": False}
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int((Path("/sys/bus/pci/devices") / bdf / "numa_node").read_text().strip())
        if node < 0:
            return info
        info["node"] = node
        cpulist = (Path("/sys/devices/system/node") / f"node{node}" / "cpulist").read_text().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
        nr = {"x86_64": 238, "aarch64": 237}.get(platform.machine())
        if nr is not None:
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            MPOL_PREFERRED = 1
            rc = libc.syscall(ctypes.c_long(nr), ctypes.c_int(MPOL_PREFERRED), ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
            info["mempolicy"] = rc == 0
    except Exception:                                      # missing sysfs entries, restricted container, ...
        pass
    return info
