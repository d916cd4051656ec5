"""Sharding a candidate batch over the GPUs of one node.

Every candidate is independent: the plan is replicated per rank (one process per GPU), rank r owns the
contiguous slice [r N / G, (r + 1) N / G) of the sweep and evaluates it with its own ``TnoPlan``.  The ONLY
exchange of the path is the final gather / min-reduce of objectives and feasibility flags (8 + 8 + 4 bytes per
candidate): z, g and J stay on the owning GPU.  ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests)
carries it; there is no other collective on the data path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "gather_objectives", "argmin_feasible", "count_feasible", "ShardedSweep"]


def shard_bounds(n_candidates: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice of rank ``rank``: [rank * N // G, (rank + 1) * N // G)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return (rank * n_candidates) // world_size, ((rank + 1) * n_candidates) // world_size


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def gather_objectives(f_local: torch.Tensor, flags_local: torch.Tensor, n_candidates: Optional[int] = None, group=None):
    """all_gather of (f, flags) in rank order.  Shards may differ by one element (shard_bounds); they are padded
    to the longest shard for the collective and trimmed afterwards.  Returns (f_all, flags_all)."""
    world, rank = _world(group)
    if world == 1:
        return f_local.clone(), flags_local.clone()
    if n_candidates is None:
        n = torch.tensor([f_local.numel()], dtype=torch.int64, device=f_local.device)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n, group=group)
        sizes = [int(s.item()) for s in sizes]
    else:
        sizes = [shard_bounds(n_candidates, world, r)[1] - shard_bounds(n_candidates, world, r)[0] for r in range(world)]
    pad = max(sizes)
    fp = torch.full((pad,), float("inf"), dtype=f_local.dtype, device=f_local.device)
    sp = torch.zeros((pad,), dtype=flags_local.dtype, device=flags_local.device)
    fp[: f_local.numel()] = f_local
    sp[: flags_local.numel()] = flags_local
    f_all = torch.empty(world * pad, dtype=f_local.dtype, device=f_local.device)
    s_all = torch.empty(world * pad, dtype=flags_local.dtype, device=flags_local.device)
    dist.all_gather_into_tensor(f_all, fp, group=group)
    dist.all_gather_into_tensor(s_all, sp, group=group)
    f_all = torch.cat([f_all[r * pad: r * pad + sizes[r]] for r in range(world)])
    s_all = torch.cat([s_all[r * pad: r * pad + sizes[r]] for r in range(world)])
    return f_all, s_all


def argmin_feasible(f_local: torch.Tensor, gmin_local: torch.Tensor, status_local: torch.Tensor, offset: int,
                    tol: float = 1e-8, group=None):
    """Best feasible candidate of the whole sweep: min f over candidates with gmin >= -tol and status == 0.
    Two all_reduce(MIN): the value, then the lowest global index attaining it (deterministic tie break).
    Returns (f_best, global_index) as Python numbers; (inf, -1) when nothing is feasible."""
    ok = (gmin_local >= -tol) & (status_local == 0) & torch.isfinite(f_local)
    fm = torch.where(ok, f_local, torch.full_like(f_local, float("inf")))
    best = fm.min().reshape(1) if fm.numel() else torch.full((1,), float("inf"), dtype=f_local.dtype, device=f_local.device)
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
    big = torch.iinfo(torch.int64).max
    if fm.numel() and torch.isfinite(best).item():
        hit = torch.nonzero(fm == best, as_tuple=False)
        idx = (hit[0, 0] + offset).reshape(1).to(torch.int64) if hit.numel() else torch.full((1,), big, dtype=torch.int64, device=f_local.device)
    else:
        idx = torch.full((1,), big, dtype=torch.int64, device=f_local.device)
    if world > 1:
        dist.all_reduce(idx, op=dist.ReduceOp.MIN, group=group)
    i = int(idx.item())
    return float(best.item()), (-1 if i == big else i)


def count_feasible(gmin_local: torch.Tensor, status_local: torch.Tensor, tol: float = 1e-8, group=None) -> int:
    c = ((gmin_local >= -tol) & (status_local == 0)).sum().to(torch.int64).reshape(1)
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return int(c.item())


class ShardedSweep:
    """One rank's view of a sweep: owns a plan on its GPU and its slice of the candidates."""

    def __init__(self, plan, n_candidates: int, group=None):
        self.plan = plan
        self.group = group
        self.world, self.rank = _world(group)
        self.n_candidates = int(n_candidates)
        self.lo, self.hi = shard_bounds(self.n_candidates, self.world, self.rank)

    def local_slice(self, x_all):
        """Rows of the full candidate matrix owned by this rank (host array or tensor)."""
        return x_all[self.lo:self.hi]

    def evaluate(self, x_local_dev, outputs=("f", "gmin", "status"), **kw):
        return self.plan.evaluate_device(x_local_dev, outputs, **kw)

    def gather(self, res):
        return gather_objectives(res["f"], res["status"], self.n_candidates, self.group)

    def best(self, res, tol: float = 1e-8):
        return argmin_feasible(res["f"], res["gmin"], res["status"], self.lo, tol, self.group)
