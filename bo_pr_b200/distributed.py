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


def broadcast_tensor(t: Tensor, src: int = 0) -> Tensor:
    """In-place broadcast of ``t`` from ``src`` (no-op in a single process)."""
    _, world = _world()
    if world > 1:
        dist.broadcast(t, src)
    return t


def broadcast_int(value: Optional[int], src: int = 0, device=None) -> Optional[int]:
    """Rank ``src``'s integer on every rank (e.g. the Sobol seed the reference takes from torch's global generator)."""
    rank, world = _world()
    if world == 1:
        return value
    buf = torch.tensor([-1 if value is None else int(value)], dtype=torch.int64, device=device)
    dist.broadcast(buf, src)
    v = int(buf.item())
    return None if v < 0 else v


def sync_pr_base_samples(acq_function, device) -> None:
    """Every rank must draw the SAME base uniforms (the reference caches one scrambled-Sobol draw per acquisition-function
    object, ``input.py:655-673``; the seed comes from torch's global generator, which differs between ranks): rank 0's
    buffers are broadcast.  No-op for acquisition functions without an MC PR transform and for the Philox stream."""
    _, world = _world()
    tf = getattr(acq_function, "input_transform", None)
    if world == 1 or tf is None or "round" not in tf:
        return
    rnd = tf["round"]
    if not hasattr(rnd, "ensure_base_samples"):
        return
    rnd.ensure_base_samples(1, device)
    for name in ("base_samples", "base_samples_categorical"):
        if hasattr(rnd, name):
            dist.broadcast(getattr(rnd, name), 0)


def allreduce_mean_(t: Tensor) -> Tensor:
    """Mean over ranks, in place (sum then divide: every rank holds an equal share of the MC samples)."""
    _, world = _world()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t.div_(world)
    return t


def mc_sharded_adam(local_value_and_grad, local_value, sync_state, ics: Tensor, lower, upper, lr: float = 0.025,
                    maxiter: int = 200) -> Tuple[Tensor, Tensor]:
    """Multi-start Adam with the MC samples -- not the restarts -- split over the ranks (fewer
    restarts than GPUs).  Every rank holds all restarts and an equal slice of the base samples; per step the local MC means
    of the value and the gradient ``[b, 1 + d]`` meet in ONE all-reduce, then every rank applies the identical Adam update
    (``gen.py:304-343`` arithmetic: clamp -> value + gradient -> Adam -> ... -> final clamp + value).

    ``local_value_and_grad(X[b, 1, d]) -> (value [b], d value.sum() / dX [b, 1, d])`` and ``local_value(X) -> [b]`` evaluate
    this rank's slice; ``sync_state()`` averages whatever estimator state depends on the MC mean (the EMA baseline is linear
    in it, so the mean of the per-rank states IS the state of the unsharded call)."""
    clamp = lambda t: torch.max(torch.min(t, upper), lower)  # noqa: E731
    theta = clamp(ics.detach().clone()).requires_grad_(True)
    opt = torch.optim.Adam([theta], lr=lr)
    b, q, d = theta.shape
    for _ in range(maxiter):
        with torch.no_grad():
            X = clamp(theta)
        v, g = local_value_and_grad(X)
        packed = torch.cat([v.reshape(b, 1), g.reshape(b, q * d)], dim=1).contiguous()
        allreduce_mean_(packed)
        sync_state()
        opt.zero_grad()
        theta.grad = -packed[:, 1:].reshape(b, q, d)
        opt.step()
    out = clamp(theta.detach())
    vals = allreduce_mean_(local_value(out).clone())
    sync_state()
    return out, vals
