"""Multi-start acquisition optimisation drivers.

API mirror of ``discrete_mixed_bo/optimize.py``: ``optimize_acqf`` (55-271) and ``optimize_acqf_mixed`` (463-589).
Every ``pr__*`` label calls ``optimize_acqf(..., stochastic=True, return_best_only=False)``
(``run_one_replication.py:494-501``): Sobol raw samples -> Boltzmann restart selection -> Adam in batches of
``options["batch_limit"]`` restarts -> per-restart values -> optional argmax.

Two extras the reference does not have, both off by default so its semantics are the default:
* ``options["batch_limit"]`` may be left out to optimise all restarts in one batch (the EMA baseline then averages over all
  of them instead of 5 at a time);
* ``shard=True`` splits the restarts over the ranks of an initialised ``torch.distributed`` group (one process per GPU, GP
  state replicated), optimises the local shard and allgathers candidates and values (``bo_pr_b200/distributed.py``).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import torch
from torch import Tensor

from .distributed import (_world, gather_restarts, mc_sharded_adam, allreduce_mean_, shard_restarts,
                          sync_pr_base_samples)
from .gen import FixedFeatureAcquisitionFunction, gen_candidates_scipy, gen_candidates_torch
from .initializers import gen_batch_initial_conditions

INIT_OPTION_KEYS = {  # keys consumed by the initialisation, not forwarded to the candidate generator (optimize.py:37-52)
    "alpha", "batch_limit", "eta", "init_batch_limit", "nonnegative", "n_burnin", "sample_around_best",
    "sample_around_best_sigma", "sample_around_best_prob_perturb", "sample_around_best_subset_sigma", "seed", "thinning",
}


def world_size() -> int:
    return _world()[1]


def _mc_shardable(acq_function, options, fixed_features) -> bool:
    from .probabilistic_reparameterization import MCProbabilisticReparameterization
    if not isinstance(acq_function, MCProbabilisticReparameterization) or fixed_features:
        return False
    rnd = acq_function.input_transform["round"]
    return rnd.rng == "sobol" and rnd.mc_samples % world_size() == 0 and not options.get("decay", False)


def _optimize_mc_sharded(pr, ics: Tensor, bounds: Tensor, options) -> Tuple[Tensor, Tensor]:
    """MC-sharded multi-start Adam on an ``MCProbabilisticReparameterization``: this rank keeps rows ``rank::world`` of the
    (already synchronised) base samples, so its forward is the MC mean over its slice; ``mc_sharded_adam`` all-reduces the
    means.  The EMA-baseline buffers are averaged after every call (the update is linear in the call's mean value)."""
    rank, world = _world()
    rnd = pr.input_transform["round"]
    full = {}
    for name in ("base_samples", "base_samples_categorical"):
        if hasattr(rnd, name):
            full[name] = getattr(rnd, name)
            rnd.register_buffer(name, full[name][rank::world].contiguous())
    mc_full = rnd.mc_samples
    rnd.mc_samples = mc_full // world
    try:
        def vg(X):
            Xg = X.detach().requires_grad_(True)
            v = pr(Xg)
            g = torch.autograd.grad(v.sum(), Xg)[0]
            return v.detach(), g

        def v_only(X):
            with torch.no_grad():
                return pr(X)

        def sync_state():
            allreduce_mean_(pr._ma_state[1:2])

        return mc_sharded_adam(vg, v_only, sync_state, ics, bounds[0], bounds[1], lr=options.get("lr", 0.025),
                               maxiter=int(options.get("maxiter", 10000)))
    finally:
        rnd.mc_samples = mc_full
        for name, t in full.items():
            rnd.register_buffer(name, t)


def optimize_acqf(
    acq_function,
    bounds: Tensor,
    q: int,
    num_restarts: int,
    raw_samples: Optional[int] = None,
    options: Optional[Dict[str, Union[bool, float, int, str]]] = None,
    inequality_constraints=None,
    equality_constraints=None,
    nonlinear_inequality_constraints=None,
    fixed_features: Optional[Dict[int, float]] = None,
    post_processing_func: Optional[Callable[[Tensor], Tensor]] = None,
    batch_initial_conditions: Optional[Tensor] = None,
    return_best_only: bool = True,
    sequential: bool = False,
    stochastic: bool = False,
    shard: bool = False,
    **kwargs: Any,
) -> Tuple[Tensor, Tensor]:
    """``optimize.py:55-271``.  Returns ``((num_restarts x) q x d candidates, (num_restarts) values)``."""
    if not (bounds.ndim == 2 and bounds.shape[0] == 2):
        raise ValueError(f"bounds should be a `2 x d` tensor, current shape: {list(bounds.shape)}.")
    if inequality_constraints or equality_constraints or nonlinear_inequality_constraints:
        raise NotImplementedError("parameter constraints are outside the PR hot path")
    if sequential and q > 1:
        if not return_best_only:
            raise NotImplementedError("`return_best_only=False` only supported for joint optimization.")
        candidate_list, acq_value_list = [], []
        base_X_pending = acq_function.X_pending
        candidates = None
        for _ in range(q):
            candidate, acq_value = optimize_acqf(
                acq_function=acq_function, bounds=bounds, q=1, num_restarts=num_restarts, raw_samples=raw_samples,
                options=options or {}, fixed_features=fixed_features, post_processing_func=post_processing_func,
                batch_initial_conditions=None, return_best_only=True, sequential=False, stochastic=stochastic,
                shard=shard)
            candidate_list.append(candidate)
            acq_value_list.append(acq_value)
            candidates = torch.cat(candidate_list, dim=-2)
            acq_function.set_X_pending(torch.cat([base_X_pending, candidates], dim=-2)
                                       if base_X_pending is not None else candidates)
        acq_function.set_X_pending(base_X_pending)
        return candidates, torch.stack(acq_value_list)

    options = options or {}
    if fixed_features is not None and len(fixed_features) == bounds.shape[-1]:
        X = torch.tensor([fixed_features[i] for i in range(bounds.shape[-1])], device=bounds.device, dtype=bounds.dtype)
        X = X.expand(q, *X.shape)
        with torch.no_grad():
            return X, acq_function(X)

    if batch_initial_conditions is None:
        if raw_samples is None:
            raise ValueError("Must specify `raw_samples` when `batch_initial_conditions` is `None`.")
        batch_initial_conditions = gen_batch_initial_conditions(
            acq_function=acq_function, bounds=bounds, q=q, num_restarts=num_restarts, raw_samples=raw_samples,
            fixed_features=fixed_features, options=options, shard=shard)
    elif shard:
        sync_pr_base_samples(acq_function, bounds.device)

    n_total = batch_initial_conditions.shape[0]
    if shard and stochastic and n_total < world_size() and _mc_shardable(acq_function, options, fixed_features):
        # fewer restarts than GPUs: split the MC samples instead, one [b, 1 + d] all-reduce per Adam step
        cand, vals = _optimize_mc_sharded(acq_function, batch_initial_conditions, bounds, options)
        if post_processing_func is not None:
            cand = post_processing_func(cand)
        if return_best_only:
            best = torch.argmax(vals.view(-1), dim=0)
            cand, vals = cand[best], vals[best]
        return cand, vals
    local_ics = shard_restarts(batch_initial_conditions) if shard else batch_initial_conditions
    batch_limit: int = options.get("batch_limit", max(local_ics.shape[0], 1))
    gen_candidates = gen_candidates_torch if stochastic else gen_candidates_scipy
    cand_list: List[Tensor] = []
    val_list: List[Tensor] = []
    for ics in local_ics.split(batch_limit):
        cand, vals = gen_candidates(
            initial_conditions=ics, acquisition_function=acq_function, lower_bounds=bounds[0], upper_bounds=bounds[1],
            options={k: v for k, v in options.items() if k not in INIT_OPTION_KEYS}, fixed_features=fixed_features)
        cand_list.append(cand)
        val_list.append(vals)
    batch_candidates = torch.cat(cand_list) if cand_list else local_ics
    batch_acq_values = torch.cat(val_list) if val_list else local_ics.new_empty(0)
    if shard:
        batch_candidates = gather_restarts(batch_candidates, n_total)
        batch_acq_values = gather_restarts(batch_acq_values, n_total)
    if post_processing_func is not None:
        batch_candidates = post_processing_func(batch_candidates)
    if return_best_only:
        best = torch.argmax(batch_acq_values.view(-1), dim=0)
        batch_candidates = batch_candidates[best]
        batch_acq_values = batch_acq_values[best]
    return batch_candidates, batch_acq_values


def optimize_acqf_mixed(
    acq_function,
    bounds: Tensor,
    q: int,
    num_restarts: int,
    fixed_features_list: List[Dict[int, float]],
    raw_samples: Optional[int] = None,
    options: Optional[Dict[str, Union[bool, float, int, str]]] = None,
    inequality_constraints=None,
    equality_constraints=None,
    post_processing_func: Optional[Callable[[Tensor], Tensor]] = None,
    batch_initial_conditions: Optional[Tensor] = None,
    **kwargs: Any,
) -> Tuple[Tensor, Tensor]:
    """``optimize.py:463-589``: optimise the free columns once per entry of ``fixed_features_list`` and return the best;
    sequential greedy for ``q > 1``.  ``stochastic=True`` may be passed through ``kwargs`` to use Adam instead of L-BFGS-B."""
    if not fixed_features_list:
        raise ValueError("fixed_features_list must be non-empty.")
    stochastic = bool(kwargs.get("stochastic", False))
    all_columns = set(range(bounds.shape[-1]))
    if q == 1:
        ff_candidate_list, ff_acq_value_list = [], []
        for fixed_features in fixed_features_list:
            ff_columns, ff_vals = list(zip(*fixed_features.items()))
            ff_acq_function = FixedFeatureAcquisitionFunction(
                acq_function=acq_function, d=bounds.shape[-1], columns=list(ff_columns),
                values=torch.tensor(ff_vals, dtype=bounds.dtype, device=bounds.device))
            keep_columns = sorted(all_columns - set(ff_columns))
            candidate, acq_value = optimize_acqf(
                acq_function=ff_acq_function, bounds=bounds[:, keep_columns], q=q, num_restarts=num_restarts,
                raw_samples=raw_samples, options=options or {}, post_processing_func=post_processing_func,
                batch_initial_conditions=batch_initial_conditions, return_best_only=True, stochastic=stochastic)
            ff_candidate_list.append(ff_acq_function._construct_X_full(X=candidate))
            ff_acq_value_list.append(acq_value)
        ff_acq_values = torch.stack(ff_acq_value_list)
        best = torch.argmax(ff_acq_values)
        return ff_candidate_list[best], ff_acq_values[best]
    base_X_pending = acq_function.X_pending
    candidates = torch.tensor([], device=bounds.device, dtype=bounds.dtype)
    for _ in range(q):
        candidate, _ = optimize_acqf_mixed(
            acq_function=acq_function, bounds=bounds, q=1, num_restarts=num_restarts, raw_samples=raw_samples,
            fixed_features_list=fixed_features_list, options=options or {}, post_processing_func=post_processing_func,
            batch_initial_conditions=batch_initial_conditions, **kwargs)
        candidates = torch.cat([candidates, candidate], dim=-2)
        acq_function.set_X_pending(torch.cat([base_X_pending, candidates], dim=-2)
                                   if base_X_pending is not None else candidates)
    acq_function.set_X_pending(base_X_pending)
    return candidates, acq_function(candidates)
