"""Trust-region bookkeeping around the PR path.

Host logic only -- a handful of scalars per BO iteration.  ``TurboState`` / ``update_state`` keep the field names, defaults
and transition rules of ``discrete_mixed_bo/trust_region.py:16-85``; ``trust_region_bounds`` restates the box computation
``run_one_replication.py:410-440`` performs before handing ``bounds`` to ``optimize_acqf`` (config ``mixed_int_f1_TR``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence

import torch
from torch import Tensor


@dataclass
class TurboState:
    """Side length and success/failure counters of one trust region (``trust_region.py:16-35``)."""

    dim: int
    batch_size: int
    is_constrained: bool
    length: float = 0.8
    length_min: float = 0.5**7
    length_max: float = 1.6
    failure_counter: int = 0
    failure_tolerance: int = float("nan")  # set from dim / batch_size below
    success_counter: int = 0
    success_tolerance: int = 10
    best_value: float = -float("inf")  # maximisation
    constraint_violation = float("inf")
    restart_triggered: bool = False

    def __post_init__(self) -> None:
        self.failure_tolerance = math.ceil(max(4.0 / self.batch_size, float(self.dim) / self.batch_size))


def _improves(new: float, old: float) -> bool:
    """Relative 1e-3 improvement test used for both the objective and the violation (``trust_region.py:40-42, 58-62``)."""
    return new > old + 1e-3 * math.fabs(old)


def update_state(state: TurboState, Y_next: Tensor) -> TurboState:
    """Advance the counters with the newest observations and resize the region (``trust_region.py:38-85``).

    Unconstrained: column 0 (or the whole tensor) is the objective.  Constrained: ``Y_next[..., 0]`` is the objective and
    ``Y_next[..., 1:] >= 0`` means feasible; the four cases below follow the reference's ordering.
    """
    if not state.is_constrained:
        y_max = Y_next.max().item()
        success = _improves(y_max, state.best_value)
        state.best_value = max(state.best_value, y_max)
    else:
        feasible = (Y_next[..., 1:] >= 0).all(dim=-1)
        any_feasible = bool(feasible.any())
        if any_feasible and state.constraint_violation == 0:  # incumbent and newcomer both feasible
            y_max = Y_next[feasible, 0].max().item()
            success = _improves(y_max, state.best_value)
            state.best_value = max(state.best_value, y_max)
        elif any_feasible and state.constraint_violation > 0:  # first feasible point
            success = True
            state.best_value = Y_next[feasible, 0].max().item()
            state.constraint_violation = 0.0
        elif not any_feasible and state.constraint_violation > 0:  # still infeasible: compare total violation
            v_min = torch.clamp_max(Y_next[..., 1:], 0.0).abs().sum(dim=-1).min().item()
            success = v_min < state.constraint_violation - 1e-3 * math.fabs(state.constraint_violation)
            state.constraint_violation = min(state.constraint_violation, v_min)
        else:  # feasible incumbent, infeasible newcomers
            success = False

    if success:
        state.success_counter += 1
        state.failure_counter = 0
    else:
        state.success_counter = 0
        state.failure_counter += 1

    if state.success_counter == state.success_tolerance:
        state.length = min(2.0 * state.length, state.length_max)
        state.success_counter = 0
    elif state.failure_counter == state.failure_tolerance:
        state.length /= 2.0
        state.failure_counter = 0

    if state.length < state.length_min:
        state.restart_triggered = True
    return state


def trust_region_center(X: Tensor, Y: Tensor, is_constrained: bool = False) -> Tensor:
    """Incumbent the region is centred on (``run_one_replication.py:413-423``): the best feasible point, else the least
    violating one; plain argmax when unconstrained."""
    if not is_constrained:
        return X[Y.argmax()].clone()
    feasible = (Y[..., 1:] >= 0).all(dim=-1)
    if feasible.any():
        merit = Y[:, 0].clone()
        merit[~feasible] = -float("inf")
        return X[merit.argmax()].clone()
    violation = torch.clamp_max(Y[..., 1:], 0.0).abs().sum(dim=-1)
    return X[violation.argmin()].clone()


def trust_region_bounds(
    state: TurboState,
    X: Tensor,
    Y: Tensor,
    standard_bounds: Tensor,
    binary_dims: Optional[Sequence[int]] = None,
    categorical_start: Optional[int] = None,
) -> Tensor:
    """``2 x d`` box ``x_center -/+ length * (ub - lb)`` clipped to ``standard_bounds`` (``run_one_replication.py:410-440``).

    Binary columns and every one-hot column from ``categorical_start`` on are reset to ``[0, 1]``: the region only
    restricts continuous and wide-integer parameters.
    """
    if X.is_cuda:
        return _trust_region_bounds_device(state, X, Y, standard_bounds, binary_dims, categorical_start)
    span = state.length * (standard_bounds[1] - standard_bounds[0])
    center = trust_region_center(X, Y, state.is_constrained).to(standard_bounds)
    bounds = torch.stack((torch.maximum(center - span, standard_bounds[0]), torch.minimum(center + span, standard_bounds[1])))
    if binary_dims is not None and len(binary_dims) > 0:
        idx = torch.as_tensor(list(binary_dims), dtype=torch.long, device=bounds.device)
        bounds[0, idx] = 0
        bounds[1, idx] = 1
    if categorical_start is not None:
        bounds[0, categorical_start:] = 0
        bounds[1, categorical_start:] = 1
    return bounds


def _trust_region_bounds_device(state: TurboState, X: Tensor, Y: Tensor, standard_bounds: Tensor, binary_dims,
                                categorical_start) -> Tensor:
    """The same box from ``prb_trust_region_bounds``: centre selection (argmax / best feasible / least violating), span,
    clipping and the [0, 1] reset of binary and one-hot columns in one kernel on the current stream -- no host read of
    ``Y.argmax()`` in the middle of a BO iteration."""
    import ctypes as C

    from . import _capi as capi

    lib = capi.load_library()
    n, d = X.shape[0], X.shape[-1]
    Xd = X.detach().to(torch.float64).reshape(n, d).contiguous()
    Yd = Y.detach().to(device=X.device, dtype=torch.float64).reshape(n, -1).contiguous()
    n_con = Yd.shape[1] - 1 if state.is_constrained else 0
    if not state.is_constrained and Yd.shape[1] != 1:
        Yd = Yd[:, :1].contiguous()
    sb = standard_bounds.detach().to(device=X.device, dtype=torch.float64).contiguous()
    reset = torch.zeros(d, dtype=torch.uint8)
    if binary_dims is not None and len(binary_dims) > 0:
        reset[torch.as_tensor(list(binary_dims), dtype=torch.long)] = 1
    if categorical_start is not None:
        reset[categorical_start:] = 1
    reset = reset.to(X.device)
    out = torch.empty(2, d, dtype=torch.float64, device=X.device)
    with torch.cuda.device(X.device):
        st = torch.cuda.current_stream().cuda_stream
        capi.check(lib.prb_trust_region_bounds(Xd.data_ptr(), Yd.data_ptr(), n, d, n_con, float(state.length),
                                               sb[0].data_ptr(), sb[1].data_ptr(), reset.data_ptr(), out.data_ptr(), None,
                                               st), "prb_trust_region_bounds")
    return out.to(standard_bounds.dtype)
