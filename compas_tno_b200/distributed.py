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

__all__ = ["bind_to_gpu_numa", "shard_bounds", "gather_objectives", "gather_packed", "best_of_packed", "argmin_feasible", "count_feasible", "ShardedSweep"]


def shard_bounds(n_candidates: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice of rank ``rank``: [rank * N // G, (rank + 1) * N // G)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return (rank * n_candidates) // world_size, ((rank + 1) * n_candidates) // world_size


def bind_to_gpu_numa(device_index: int):
    """Pin this process (and the threads it creates later: the host fill threads of tno_fetch_jacobian_host) to the CPUs
    of the NUMA node the GPU hangs off, so that pinned host buffers allocated afterwards are first-touched on that node
    and the ranks of one box do not all write into one socket's memory.  Linux sysfs only; returns a small dict saying
    what was done (``bound`` False when the topology is not exposed, e.g. numa_node = -1 inside a container)."""
    import os
    info = {"bound": False}
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        base = "/sys/bus/pci/devices/" + bus
        node = int(open(base + "/numa_node").read().strip())
        info.update(pci=bus, numa_node=node)
        if node < 0:
            return info
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as exc:   # topology not exposed: stay unbound, say why
        info["error"] = "%s: %s" % (type(exc).__name__, exc)
    return info


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


def gather_packed(f_local: torch.Tensor, gmin_local: torch.Tensor, status_local: torch.Tensor, n_candidates: int, group=None,
                  out: Optional[torch.Tensor] = None):
    """ONE collective per sweep: the three per-candidate results the sweep exchanges (objective, feasibility margin,
    status word: 24 bytes per candidate) are packed into one (n_local, 3) float64 block and all-gathered.  Every rank
    then holds the whole sweep's (f, gmin, status) in candidate order and can take the best feasible candidate locally,
    with no further reduction.  Shards may differ by one element (shard_bounds); they are padded to the longest one
    (f = +inf, status = -1 in the padding) and trimmed afterwards.  Returns the (n_candidates, 3) float64 tensor."""
    world, rank = _world(group)
    packed = torch.stack([f_local, gmin_local, status_local.to(f_local.dtype)], dim=1)
    if world == 1:
        return packed
    sizes = [shard_bounds(n_candidates, world, r)[1] - shard_bounds(n_candidates, world, r)[0] for r in range(world)]
    pad = max(sizes)
    if packed.shape[0] != pad:
        fill = torch.tensor([float("inf"), float("-inf"), -1.0], dtype=packed.dtype, device=packed.device)
        packed = torch.cat([packed, fill.expand(pad - packed.shape[0], 3)], dim=0)
    if out is None or tuple(out.shape) != (world * pad, 3):
        out = torch.empty((world * pad, 3), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed.contiguous(), group=group)
    if all(sz == pad for sz in sizes):
        return out
    return torch.cat([out[r * pad: r * pad + sizes[r]] for r in range(world)], dim=0)


def best_of_packed(packed: torch.Tensor, tol: float = 1e-8):
    """(f_best, global index, number feasible) from the gathered (N, 3) block of gather_packed; (inf, -1, 0) when nothing
    is feasible.  Lowest index wins ties, like argmin_feasible.  One host read of three numbers."""
    f, gmin, status = packed[:, 0], packed[:, 1], packed[:, 2]
    ok = (gmin >= -tol) & (status == 0) & torch.isfinite(f)
    fm = torch.where(ok, f, torch.full_like(f, float("inf")))
    if fm.numel() == 0:
        return float("inf"), -1, 0
    idx = torch.argmin(fm)          # first minimal element
    res = torch.stack([fm[idx], idx.to(fm.dtype), ok.sum().to(fm.dtype)]).cpu()
    best, i, cnt = float(res[0]), int(res[1]), int(res[2])
    if not (best < float("inf")):
        return float("inf"), -1, 0
    # torch.argmin does not promise the first of equal minima on every backend: settle ties explicitly
    first = int(torch.nonzero(fm == fm[idx], as_tuple=False)[0, 0])
    return best, min(i, first), cnt


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

    def gather_packed(self, res, out=None):
        """One all_gather of (f, gmin, status) for the whole sweep; see gather_packed."""
        return gather_packed(res["f"], res["gmin"], res["status"], self.n_candidates, self.group, out)
