"""Host-side sharding arithmetic shared by bench.py and the tests.

The engine (gomcb200_set_shard) gives rank r of W the contiguous share
[n*r/W, n*(r+1)/W) of an n-unit work line (cells for the pair sweep, weighted
(tile, chunk) units for the structure factor); the helpers below restate that
split for host code and reduce the three partial energies."""
from __future__ import annotations


def split_range(n: int, rank: int, world: int):
    """Contiguous share of n units for `rank` (same integer arithmetic as engine.cu)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def allreduce_energies(values, world: int, device=None):
    """Sum the partial (LJ, real, recip) energies over ranks; identity for one rank."""
    if world == 1:
        return list(values)
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t)
    return t.tolist()


def max_over_ranks(ms: float, world: int, device=None) -> float:
    """Multi-GPU timings are the maximum over ranks."""
    if world == 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
