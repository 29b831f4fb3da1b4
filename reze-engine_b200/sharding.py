"""Multi-GPU: a crowd of independent character instances partitions across GPUs.

One process per GPU (torchrun).  Rank r of W owns the contiguous instance range `instance_range(K, W, r)`; the mesh,
morph and SDEF tables are replicated by each rank's own `rz_load_*` calls, palettes / morph weights are fed per rank, outputs
stay on the owning GPU.  There is NO collective on the data path; the only exchange is the trivial end-of-frame gather
of one small record per GPU (`gather_records`), which works on any torch.distributed backend (NCCL on GPUs, gloo in the
CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def instance_range(K_total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[first, last) of the instances rank `rank` owns; ranges are contiguous, disjoint, cover 0..K_total, and differ in
    size by at most one instance."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(K_total, world_size)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def owner_of(instance: int, K_total: int, world_size: int) -> int:
    base, rem = divmod(K_total, world_size)
    cut = rem * (base + 1)
    if instance < cut:
        return instance // (base + 1)
    return rem + (instance - cut) // max(base, 1)


def gather_records(record: Sequence[float], group=None) -> List[List[float]]:
    """All ranks contribute one small float record (verts done, device ms, checksum, ...); every rank gets all of them.
    This is the "trivial result gather": <= 64 bytes per GPU, once per frame at most."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [list(map(float, record))]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor(list(record), dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return [[float(x) for x in o.tolist()] for o in out]


def max_over_ranks(value: float, group=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device_index: int) -> dict:
    """Best effort: restrict this process to the CPUs of the NUMA node GPU `device_index` hangs off, BEFORE any pinned
    allocation is made, so that the rank's staging buffers (first touched by these threads) and its launch thread sit next
    to its GPU — with 8 ranks per host the per-frame palette uploads then do not all cross the socket interconnect.
    Returns what was done (for the bench record); never raises."""
    import os
    info = {"bound": False}
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if vis:
            parts = vis.split(",")
            if device_index < len(parts) and parts[device_index].strip().isdigit():
                idx = int(parts[device_index])
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        sysdir = "/sys/bus/pci/devices/" + bus.lower()[-12:]                 # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(sysdir + "/numa_node").read())
        cpus = _parse_cpulist(open(sysdir + "/local_cpulist").read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        info.update(numa_node=node, cpus=len(allowed))
        if node >= 0 and allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
    except Exception as e:  # noqa: BLE001
        info["error"] = repr(e)[:120]
    return info
