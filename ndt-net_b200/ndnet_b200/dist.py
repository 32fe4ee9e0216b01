"""Host-side logic of the multi-GPU path: scans are independent units, so rank r simply owns a disjoint slice of
them (no collective on the forward path).  The only communication is the barrier around a timed
region and the max-over-ranks of its duration; both work with NCCL (GPU tensors) and gloo (CPU tensors)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced slice of `n_items` units owned by `rank` (first n % world ranks get one more)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def scan_seeds(rank: int, batch: int, set_index: int, per_rank_stride: int = 100_000) -> range:
    """Seeds of the synthetic scans rank `rank` generates for resident input set `set_index` (disjoint over ranks)."""
    start = rank * per_rank_stride + set_index * batch
    return range(start, start + batch)


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> str:
    """Best effort: restrict this process to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host buffers
    are allocated (first touch then places them on that node), so that the host<->device copies of the e2e path do not
    cross the socket interconnect when several ranks share the box.  Returns a one-line description; never raises."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id if hasattr(
            torch.cuda.get_device_properties(device_index), "pci_bus_id") else None
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        if bus is None:
            return "numa: no pci id"
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa: single node"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        mine = os.sched_getaffinity(0) & cpus
        if not mine:
            return f"numa: node {node} has none of this process's CPUs"
        os.sched_setaffinity(0, mine)
        return f"numa: bound to node {node} ({len(mine)} cpus)"
    except Exception as exc:                      # sysfs layout, permissions: keep going unbound
        return f"numa: not bound ({type(exc).__name__})"
