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
