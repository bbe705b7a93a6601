"""Restart initialisation for multi-start acquisition optimisation.

API mirror of ``discrete_mixed_bo/initializers.py`` for the PR path: ``gen_batch_initial_conditions`` (53-206),
``initialize_q_batch`` (462-534), ``initialize_q_batch_nonneg`` (537-606).  Raw samples are scrambled-Sobol points drawn
on the host (as the reference does, :138), screened through the acquisition function on the device in chunks of
``init_batch_limit`` -- the chunking is kept because every PR forward advances the EMA baseline, so
the number of screening calls is observable -- and restarts are selected by the Boltzmann heuristic with torch's global
generator.  Constraints, ``sample_around_best`` and one-shot acquisition functions are outside the PR path (raise).
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Optional, Tuple, Union

import torch
from torch import Tensor
from torch.quasirandom import SobolEngine

from .acquisition import is_nonnegative


class BadInitialCandidatesWarning(RuntimeWarning):
    """``botorch.exceptions.BadInitialCandidatesWarning``."""


def draw_sobol_samples(bounds: Tensor, n: int, q: int, seed: Optional[int] = None) -> Tensor:
    """``botorch.utils.sampling.draw_sobol_samples``: ``[n, q, d]`` scrambled Sobol in ``bounds``."""
    d = bounds.shape[-1]
    u = SobolEngine(q * d, scramble=True, seed=seed).draw(n, dtype=bounds.dtype).view(n, q, d)
    return bounds[0] + (bounds[1] - bounds[0]) * u


def fix_features(X: Tensor, fixed_features: Optional[Dict[int, Optional[float]]] = None) -> Tensor:
    """``botorch.optim.utils.fix_features``: overwrite the given columns (``None`` value: detach that column)."""
    if fixed_features is None:
        return X
    cols = []
    for i in range(X.shape[-1]):
        if i in fixed_features:
            v = fixed_features[i]
            cols.append(X[..., i].detach() if v is None else torch.full_like(X[..., i], float(v)))
        else:
            cols.append(X[..., i])
    return torch.stack(cols, dim=-1)


def initialize_q_batch(X: Tensor, Y: Tensor, n: int, eta: float = 1.0) -> Tensor:
    """``initializers.py:462-534``: sample ``n`` rows of ``X`` with probability proportional to ``exp(eta * zscore(Y))``,
    always keeping the arg-max (q-batch shape ``b x q x d`` only; t-batches of initial conditions are not used by PR)."""
    n_samples = X.shape[0]
    if X.ndim != 3:
        raise NotImplementedError("batched initial conditions are not supported")
    if n > n_samples:
        raise RuntimeError(f"n ({n}) cannot be larger than the number of provided samples ({n_samples})")
    if n == n_samples:
        return X
    Ystd = Y.std(dim=0)
    if bool(torch.any(Ystd == 0)):
        warnings.warn("All acquisition values for raw samples points are the same for at least one batch. "
                      "Choosing initial conditions at random.", BadInitialCandidatesWarning)
        return X[torch.randperm(n=n_samples, device=X.device)][:n]
    max_idx = torch.argmax(Y, dim=0)
    etaZ = eta * (Y - Y.mean(dim=0)) / Ystd
    weights = torch.exp(etaZ)
    while bool(torch.isinf(weights).any()):
        etaZ = etaZ * 0.5
        weights = torch.exp(etaZ)
    idcs = torch.multinomial(weights, n)
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return X[idcs]


def initialize_q_batch_nonneg(X: Tensor, Y: Tensor, n: int, eta: float = 1.0, alpha: float = 1e-4) -> Tensor:
    """``initializers.py:537-606``: like ``initialize_q_batch`` for non-negative acquisition functions (EI): ignore
    samples with ``Y < alpha * max(Y)``, weights ``exp(eta (Y / max(Y) - 1))``."""
    n_samples = X.shape[0]
    if n > n_samples:
        raise RuntimeError("n cannot be larger than the number of provided samples")
    if n == n_samples:
        return X
    max_val, max_idx = torch.max(Y, dim=0)
    if bool(torch.any(max_val <= 0)):
        warnings.warn("All acquisition values for raw sampled points are nonpositive, so initial conditions are being "
                      "selected randomly.", BadInitialCandidatesWarning)
        return X[torch.randperm(n=n_samples, device=X.device)][:n]
    pos = Y > 0
    num_pos = int(pos.sum())
    if num_pos < n:
        remaining = (~pos).nonzero(as_tuple=False).view(-1)
        pick = remaining[torch.randperm(remaining.shape[0], device=Y.device)[: n - num_pos]]
        pos[pick] = True
        return X[pos]
    alpha_pos = Y >= alpha * max_val
    while int(alpha_pos.sum()) < n:
        alpha = 0.1 * alpha
        alpha_pos = Y >= alpha * max_val
    candidates = torch.arange(len(Y), device=Y.device)[alpha_pos]
    weights = torch.exp(eta * (Y[alpha_pos] / max_val - 1))
    idcs = candidates[torch.multinomial(weights, n)]
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return X[idcs]


def gen_batch_initial_conditions(
    acq_function,
    bounds: Tensor,
    q: int,
    num_restarts: int,
    raw_samples: int,
    fixed_features: Optional[Dict[int, float]] = None,
    options: Optional[Dict[str, Union[bool, float, int]]] = None,
    inequality_constraints: Optional[List[Tuple[Tensor, Tensor, float]]] = None,
    equality_constraints: Optional[List[Tuple[Tensor, Tensor, float]]] = None,
    shard: bool = False,
) -> Tensor:
    """``initializers.py:53-206``.  Returns ``num_restarts x q x d`` initial conditions on ``bounds.device``.

    ``shard=True`` (inside an initialised ``torch.distributed`` group): the raw samples are drawn from ONE seed (rank 0's,
    broadcast), every rank screens the rows ``rank::world`` (``initializers.py:179-191`` is the loop being split), the
    values are all-gathered back into raw-sample order, rank 0 runs the Boltzmann selection with its generator and the
    chosen initial conditions are broadcast -- every rank returns the same tensor.  With ``init_batch_limit`` each rank
    makes ``1/world`` of the screening calls, so the EMA baseline of a ``reinforce_ma`` PR function has seen fewer updates
    when the optimisation starts than in the single-process run (that state is per-call by design)."""
    if inequality_constraints is not None or equality_constraints is not None:
        raise NotImplementedError("parameter constraints are outside the PR hot path")
    options = options or {}
    if options.get("sample_around_best", False):
        raise NotImplementedError("sample_around_best is not used by the pr__* labels")
    seed: Optional[int] = options.get("seed")
    batch_limit: Optional[int] = options.get("init_batch_limit", options.get("batch_limit"))
    init_kwargs = {}
    if "eta" in options:
        init_kwargs["eta"] = options.get("eta")
    if options.get("nonnegative") or is_nonnegative(acq_function):
        init_func = initialize_q_batch_nonneg
        if "alpha" in options:
            init_kwargs["alpha"] = options.get("alpha")
    else:
        init_func = initialize_q_batch
    q = 1 if q is None else q
    device = bounds.device
    bounds_cpu = bounds.cpu()
    factor, max_factor = 1, 5
    batch_initial_conditions = None
    rank, world = (0, 1)
    if shard:
        from . import distributed as D
        rank, world = D._world()
        if world > 1:
            if seed is None:  # the reference lets SobolEngine seed itself; ranks must agree, so rank 0 picks
                seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
            seed = D.broadcast_int(seed, 0, device=device)
            D.sync_pr_base_samples(acq_function, device)
    while factor < max_factor:
        with warnings.catch_warnings(record=True) as ws:
            warnings.simplefilter("always")
            n = raw_samples * factor
            X_rnd = draw_sobol_samples(bounds=bounds_cpu, n=n, q=q, seed=seed)
            X_rnd = fix_features(X_rnd, fixed_features=fixed_features)
            with torch.no_grad():
                limit = X_rnd.shape[0] if batch_limit is None else batch_limit
                # one host->device copy and one device->host read for the whole screening pass (the reference moves every
                # chunk separately, initializers.py:184-190; the chunked CALLS are kept, they advance the EMA state)
                X_dev = X_rnd.to(device=device)
                X_loc = X_dev[rank::world] if world > 1 else X_dev
                if hasattr(acq_function, "screen") and q == 1:
                    # MC-PR acquisition functions run the chunk loop inside the library (prb_mc_screen)
                    Y_loc = acq_function.screen(X_loc, limit)
                else:
                    Y_loc = torch.cat([acq_function(X_loc[s:s + limit]) for s in range(0, X_loc.shape[0], limit)])
                if world > 1:
                    Y_loc = D.gather_restarts(Y_loc.contiguous(), X_dev.shape[0])
                Y_rnd = Y_loc.cpu()
            bad = False
            if rank == 0:
                batch_initial_conditions = init_func(X=X_rnd, Y=Y_rnd, n=num_restarts, **init_kwargs).to(device=device)
                bad = any(issubclass(w.category, BadInitialCandidatesWarning) for w in ws)
            if world > 1:
                if rank != 0:
                    batch_initial_conditions = torch.empty(num_restarts, q, bounds.shape[-1], dtype=X_rnd.dtype,
                                                           device=device)
                D.broadcast_tensor(batch_initial_conditions, 0)
                bad = bool(D.broadcast_int(int(bad), 0, device=device))
            if not bad:
                return batch_initial_conditions
            factor += 1
            if seed is not None:
                seed += 1
    warnings.warn("Unable to find non-zero acquisition function values - initial conditions are being selected "
                  "randomly.", BadInitialCandidatesWarning)
    return batch_initial_conditions
