"""Multi-GPU plumbing (one process per GPU).  Inference shards by SONG with no data-path collective
(songs and even patches are independent: reference data.py:66, inference.py:63,79-116); training is data
parallel with one gradient all-reduce per step.  Only timing / bookkeeping reductions live here."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_songs(n_songs: int, rank: int, world: int) -> list[int]:
    """Round-robin: song i goes to rank i % world."""
    return list(range(rank, n_songs, world))


def _reduce(value: float, op, backend_device: str) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=backend_device)
    dist.all_reduce(t, op=op)
    return float(t[0])


def max_over_ranks(value: float, backend_device: str = "cuda") -> float:
    return _reduce(value, dist.ReduceOp.MAX, backend_device)


def sum_over_ranks(value: float, backend_device: str = "cuda") -> float:
    return _reduce(value, dist.ReduceOp.SUM, backend_device)


def average_flat_(flat: torch.Tensor) -> torch.Tensor:
    """In-place mean of a flat gradient buffer across ranks (single collective)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat)
        flat.div_(dist.get_world_size())
    return flat
