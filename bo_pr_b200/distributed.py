"""Multi-GPU plumbing of the PR path: shard restarts / raw samples across ranks, replicate the GP state, and do the ONE
exchange of the path -- the best-candidate allgather + argmax (``discrete_mixed_bo/optimize.py:262-265``;
``run_one_replication.py:523-528``).  One process per GPU over ``torch.distributed`` (NCCL on the box, gloo in the CPU
tests).  No collective runs inside an optimisation step: every restart's MC reduction is local to one rank
(SURVEY.md section 8e)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_slice(n_items: int, rank: Optional[int] = None, world: Optional[int] = None) -> slice:
    """Round-robin shard ``r -> rank r mod G`` (restart r goes to GPU r mod G), as a strided slice."""
    r, w = _world()
    rank = r if rank is None else rank
    world = w if world is None else world
    return slice(rank, n_items, world)


def shard_restarts(X: Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> Tensor:
    """This rank's restarts (or raw samples) of ``X [r, q, d]``."""
    return X[shard_slice(X.shape[0], rank, world)]


def gather_restarts(local: Tensor, n_total: int) -> Tensor:
    """Inverse of ``shard_restarts``: allgather the per-rank rows back into global restart order.  ``local [r_local, ...]``
    (ranks may hold one row fewer than rank 0 when ``n_total`` is not a multiple of the world size)."""
    rank, world = _world()
    if world == 1:
        return local
    per = (n_total + world - 1) // world
    pad = torch.zeros(per, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out = torch.empty(n_total, *local.shape[1:], dtype=local.dtype, device=local.device)
    for r in range(world):
        cnt = len(range(r, n_total, world))
        out[r:n_total:world] = bufs[r][:cnt]
    return out


def gather_best(values: Tensor, candidates: Tensor) -> Tuple[Tensor, Tensor, int]:
    """argmax over every rank's restarts: each rank contributes its local best ``(value, candidate[d])``; returns
    ``(best value, best candidate, owning rank)`` on all ranks.  Ties resolve to the lowest rank (deterministic)."""
    rank, world = _world()
    i = int(torch.argmax(values))
    packed = torch.cat([values[i].reshape(1), candidates[i].reshape(-1)]).contiguous()
    if world == 1:
        return packed[0], packed[1:].reshape(candidates.shape[1:]), 0
    bufs = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(bufs, packed)
    allv = torch.stack([b[0] for b in bufs])
    r = int(torch.argmax(allv))  # first maximum => lowest rank
    return bufs[r][0], bufs[r][1:].reshape(candidates.shape[1:]), r


def broadcast_gp_data(train_X: Tensor, train_Y: Tensor, src: int = 0) -> Tuple[Tensor, Tensor]:
    """Replicate the GP training data from ``src`` (every rank then builds the identical device GP state)."""
    rank, world = _world()
    if world > 1:
        dist.broadcast(train_X, src)
        dist.broadcast(train_Y, src)
    return train_X, train_Y
