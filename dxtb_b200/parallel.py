"""Multi-GPU sharding of a molecule batch: one process per GPU, no collective on the data path.

Molecules of a batch are independent in dxtb's default SCF (per-system Anderson mixer and convergence,
``scf/mixer/anderson.py:249-266``, ``scf/unrolling/default.py:213-321``), so every rank builds its own
``GFN1Calculator`` for a contiguous shard and only the final results are gathered.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced [start, stop) of ``n`` molecules for ``rank`` (first ``n % world`` ranks get one more)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("invalid rank / world_size")
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_by_cost(cost: torch.Tensor, world_size: int) -> list[torch.Tensor]:
    """Greedy longest-processing-time assignment for ragged batches (cost ~ nao^3): returns per-rank index tensors."""
    order = torch.argsort(cost, descending=True).tolist()
    loads = [0.0] * world_size
    out: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=loads.__getitem__)
        out[r].append(i)
        loads[r] += float(cost[i])
    return [torch.tensor(sorted(ix), dtype=torch.long) for ix in out]


def gather_results(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-molecule results (leading dim = local molecules) of contiguous shards into (n_total, ...).
    This is the only communication of a sharded single point: O(nb) numbers, after the kernels have finished."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    nmax = max(b - a for a, b in sizes)
    pad = local.new_zeros((nmax, *local.shape[1:]))
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: b - a] for r, (a, b) in enumerate(sizes)], dim=0)


def gather_by_index(local: torch.Tensor, parts, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-molecule results of arbitrary (e.g. cost-balanced) shards: ``parts[r]`` holds the global molecule
    ids of rank ``r`` in the order of its local batch.  Returns (n_total, ...) in global order on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    nmax = max(len(p) for p in parts)
    pad = local.new_zeros((nmax, *local.shape[1:]))
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = local.new_zeros((n_total, *local.shape[1:]))
    for r in range(world):
        idx = torch.as_tensor(parts[r], dtype=torch.long, device=local.device)
        out[idx] = bufs[r][: idx.numel()]
    return out
