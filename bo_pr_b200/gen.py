"""Candidate generation: multi-start Adam (``gen_candidates_torch``) and L-BFGS-B (``gen_candidates_scipy``).

API mirror of ``discrete_mixed_bo/gen.py``.  ``gen_candidates_torch`` (264-343) is what every ``pr__*`` label uses
(``stochastic=True``): clamp -> value + gradient -> Adam(lr 0.025) for exactly ``maxiter`` steps when ``rel_tol`` is unset
-> final clamp and value.  When the acquisition function is an ``MCProbabilisticReparameterization`` and nothing needs the
host in the loop (no callback, no decay, plain Adam), the whole loop runs on the device in ``prb_mc_adam``: one captured
CUDA graph per step, all restarts of the batch at once -- that is the reference's 200 host round trips per restart batch
collapsed into one call.  Otherwise the loop is host-driven with the same arithmetic (torch.optim.Adam).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, NoReturn, Optional, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from . import _capi as capi


def columnwise_clamp(X: Tensor, lower=None, upper=None, raise_on_violation: bool = False) -> Tensor:
    """``botorch.optim.utils.columnwise_clamp``: elementwise max/min with per-column (or scalar) bounds."""
    out = X
    if lower is not None:
        out = torch.max(out, torch.as_tensor(lower, dtype=X.dtype, device=X.device).expand_as(out))
    if upper is not None:
        out = torch.min(out, torch.as_tensor(upper, dtype=X.dtype, device=X.device).expand_as(out))
    if raise_on_violation and not torch.equal(X, out):
        raise RuntimeError("Original value(s) are out of bounds.")
    return out


class ExpMAStoppingCriterion:
    """``botorch.optim.stopping.ExpMAStoppingCriterion`` ([3P]): stop after ``maxiter`` evaluations or
    when the relative decrease of the exponentially weighted moving average of the loss over ``n_window`` steps falls below
    ``rel_tol`` (``-inf`` turns this off, as ``gen.py:311-313`` does for PR)."""

    def __init__(self, maxiter: int = 10000, minimize: bool = True, n_window: int = 10, eta: float = 1.0,
                 rel_tol: float = 1e-5):
        self.maxiter, self.minimize, self.n_window, self.rel_tol = maxiter, minimize, n_window, rel_tol
        self.iter = 0
        w = torch.exp(torch.linspace(-eta, 0, n_window))
        self.weights = w / w.sum()
        self._prev_fvals = None

    def evaluate(self, fvals: Tensor) -> bool:
        self.iter += 1
        if self.iter == self.maxiter:
            return True
        f = fvals.detach().reshape(1, -1).cpu().double()
        self._prev_fvals = f if self._prev_fvals is None else torch.cat([self._prev_fvals[-self.n_window:], f])
        if self._prev_fvals.size(0) < self.n_window + 1:
            return False
        w = self.weights.to(f).unsqueeze(-1)
        prev_ma = (self._prev_fvals[:-1] * w).sum(dim=0)
        ma = (self._prev_fvals[1:] * w).sum(dim=0)
        rel_delta = (prev_ma - ma) / prev_ma.abs()
        if not self.minimize:
            rel_delta = -rel_delta
        return bool(torch.max(rel_delta) < self.rel_tol)


class FixedFeatureAcquisitionFunction(torch.nn.Module):
    """``botorch.acquisition.fixed_feature.FixedFeatureAcquisitionFunction``: evaluates the wrapped AF
    with some columns pinned; ``forward`` takes the free columns only."""

    def __init__(self, acq_function, d: int, columns, values):
        super().__init__()
        self.acq_func = acq_function
        self.d = d
        self._columns = sorted(int(c) for c in columns)
        order = np.argsort([int(c) for c in columns])
        vals = torch.as_tensor(values, dtype=torch.float64).reshape(-1)[torch.as_tensor(order)]
        self.register_buffer("values", vals)
        self._free = [i for i in range(d) if i not in set(self._columns)]
        self.X_pending = None

    def set_X_pending(self, X_pending=None):
        self.X_pending = X_pending

    def _construct_X_full(self, X: Tensor) -> Tensor:
        cols = []
        fi = 0
        vi = 0
        for i in range(self.d):
            if vi < len(self._columns) and self._columns[vi] == i:
                cols.append(self.values[vi].to(X).expand(X.shape[:-1]))
                vi += 1
            else:
                cols.append(X[..., fi])
                fi += 1
        return torch.stack(cols, dim=-1)

    def forward(self, X: Tensor) -> Tensor:
        return self.acq_func(self._construct_X_full(X))


def _is_fusable(acquisition_function, options, callback, optimizer, fixed_features) -> bool:
    from .probabilistic_reparameterization import (AnalyticProbabilisticReparameterization,
                                                   MCProbabilisticReparameterization)

    if not isinstance(acquisition_function, (MCProbabilisticReparameterization, AnalyticProbabilisticReparameterization)):
        return False
    if hasattr(acquisition_function.acq_function, "multi_desc"):
        return False  # model lists: host-driven loop over prb_mc_value_grad_multi
    rnd = acquisition_function.input_transform["round"]
    if getattr(rnd, "resample", False):
        return False  # fresh base samples on every forward: the device loop keeps one draw for all steps
    return (callback is None and optimizer is torch.optim.Adam and not options.get("decay", False) and not fixed_features
            and options.get("fused", True)
            and (options.get("rel_tol") is None or options.get("rel_tol") == float("-inf")))


def _fused_analytic_adam(ics: Tensor, acqf, lower, upper, lr: float, maxiter: int) -> Tuple[Tensor, Tensor]:
    """``prb_analytic_adam``: the whole multi-start Adam on the analytic PR objective in one call."""
    from .acquisition import acq_desc_of
    from .models import gp_state_from_model

    capi.require_cuda(ics, "initial_conditions")
    dev = ics.device
    b, q, d = ics.shape
    if q != 1:
        raise NotImplementedError("q > 1 is not supported by the PR path")
    state = gp_state_from_model(acqf.acq_function.model)
    codes = acqf.input_transform["round"].option_codes
    if codes.device != dev:
        codes = codes.to(dev)
    n_opt = codes.shape[0]
    theta = ics.detach().to(torch.float64).reshape(b, d).contiguous().clone()
    lo = torch.as_tensor(lower, dtype=torch.float64, device=dev).expand(d).contiguous()
    hi = torch.as_tensor(upper, dtype=torch.float64, device=dev).expand(d).contiguous()
    adam = torch.zeros(2 * b * d, dtype=torch.float64, device=dev)
    value = torch.empty(b, dtype=torch.float64, device=dev)
    ws = state.workspace(b * n_opt, d, b)
    sp, aq = acqf._space(), acq_desc_of(acqf.acq_function)
    with torch.cuda.device(dev):
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            capi.check(state.lib.prb_analytic_adam(state.handle, C.byref(sp), C.byref(aq), theta.data_ptr(), b,
                                                   codes.data_ptr(), n_opt, lo.data_ptr(), hi.data_ptr(), float(lr),
                                                   int(maxiter), 0, adam.data_ptr(), value.data_ptr(), ws.data_ptr(),
                                                   ws.numel(), int(state.chunk_cols), side.cuda_stream),
                       "prb_analytic_adam")
        torch.cuda.current_stream(dev).wait_stream(side)
        for t in (theta, lo, hi, adam, value, ws, codes):
            t.record_stream(side)
    return theta.reshape(b, 1, d).to(ics.dtype), value


def _fused_mc_adam(ics: Tensor, acqf, lower, upper, lr: float, maxiter: int) -> Tuple[Tensor, Tensor]:
    """All restarts of the batch, all ``maxiter`` Adam steps and the final evaluation in one ``prb_mc_adam`` call."""
    from .acquisition import acq_desc_of
    from .models import gp_state_from_model

    capi.require_cuda(ics, "initial_conditions")
    dev = ics.device
    b, q, d = ics.shape
    if q != 1:
        raise NotImplementedError("q > 1 is not supported by the PR path")
    state = gp_state_from_model(acqf.acq_function.model)
    state.set_option(capi.PRB_OPT_DEDUP, int(getattr(acqf, "dedup", False)))
    rnd = acqf.input_transform["round"]
    rnd.ensure_base_samples(1, dev)
    u_int, u_cat = rnd.base_sample_ptrs()
    theta = ics.detach().to(torch.float64).reshape(b, d).contiguous().clone()
    lo = torch.as_tensor(lower, dtype=torch.float64, device=dev).expand(d).contiguous()
    hi = torch.as_tensor(upper, dtype=torch.float64, device=dev).expand(d).contiguous()
    adam = torch.zeros(2 * b * d, dtype=torch.float64, device=dev)
    value = torch.empty(b, dtype=torch.float64, device=dev)
    ws = state.workspace(b * rnd.mc_samples, d, b)
    sp, aq = acqf._space(), acq_desc_of(acqf.acq_function)
    est = capi.PRB_EST_REINFORCE_MA if acqf.grad_estimator == "reinforce_ma" else capi.PRB_EST_REINFORCE
    with torch.cuda.device(dev):
        # graph capture needs a non-default stream; hand the side stream the current stream's work and wait for it after
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            capi.check(state.lib.prb_mc_adam(state.handle, C.byref(sp), C.byref(aq), theta.data_ptr(), b, rnd.mc_samples,
                                             capi.ptr(u_int), capi.ptr(u_cat), 0, 0, est, acqf.ma_counter.data_ptr(),
                                             lo.data_ptr(), hi.data_ptr(), float(lr), int(maxiter), 0, adam.data_ptr(),
                                             value.data_ptr(), ws.data_ptr(), ws.numel(), int(state.chunk_cols),
                                             side.cuda_stream), "prb_mc_adam")
        torch.cuda.current_stream(dev).wait_stream(side)
        for t in (theta, lo, hi, adam, value, ws):
            t.record_stream(side)
    return theta.reshape(b, 1, d).to(ics.dtype), value


def gen_candidates_torch(
    initial_conditions: Tensor,
    acquisition_function,
    lower_bounds: Optional[Union[float, Tensor]] = None,
    upper_bounds: Optional[Union[float, Tensor]] = None,
    options: Optional[Dict[str, Union[float, str]]] = None,
    callback: Optional[Callable[[int, Tensor, Tensor], NoReturn]] = None,
    fixed_features: Optional[Dict[int, Optional[float]]] = None,
    optimizer=torch.optim.Adam,
) -> Tuple[Tensor, Tensor]:
    """``gen.py:264-343``.  Returns ``(clamped candidates [b, q, d], acquisition values [b])``."""
    options = options or {}
    if fixed_features:
        d = initial_conditions.shape[-1]
        if any(v is None for v in fixed_features.values()):
            raise NotImplementedError("fixed_features with None values are not supported")
        cols = sorted(fixed_features)
        free = [i for i in range(d) if i not in fixed_features]
        sub_acqf = FixedFeatureAcquisitionFunction(acquisition_function, d, cols, [fixed_features[c] for c in cols])

        def _sub(bnd):
            if bnd is None or not torch.is_tensor(bnd) or bnd.ndim == 0:
                return bnd
            return bnd[..., free]

        cand, vals = gen_candidates_torch(initial_conditions[..., free], sub_acqf, _sub(lower_bounds),
                                          _sub(upper_bounds), options=options, callback=callback, fixed_features=None,
                                          optimizer=optimizer)
        return sub_acqf._construct_X_full(cand), vals

    lr = options.get("lr", 0.025)
    if callback is None:
        callback = options.get("callback")
    if options.get("rel_tol") is None:
        options["rel_tol"] = float("-inf")  # the reference writes this into the caller's dict too (gen.py:311-313)
    maxiter = int(options.get("maxiter", 10000))
    # the device loop clamps against explicit bounds; `None` bounds (the reference then does not clamp) take the host loop
    if lower_bounds is not None and upper_bounds is not None and \
            _is_fusable(acquisition_function, options, callback, optimizer, fixed_features):
        from .probabilistic_reparameterization import AnalyticProbabilisticReparameterization
        fused = (_fused_analytic_adam if isinstance(acquisition_function, AnalyticProbabilisticReparameterization)
                 else _fused_mc_adam)
        return fused(initial_conditions, acquisition_function, lower_bounds, upper_bounds, lr, maxiter)

    clamp = lambda t: columnwise_clamp(t, lower_bounds, upper_bounds)  # noqa: E731
    theta = clamp(initial_conditions).detach().clone().requires_grad_(True)
    opt = optimizer(params=[theta], lr=lr)
    stop_kwargs = {k: options[k] for k in ("maxiter", "minimize", "n_window", "eta", "rel_tol") if k in options}
    stopping = ExpMAStoppingCriterion(**stop_kwargs)
    decay = options.get("decay", False)
    i, stop = 0, False
    while not stop:
        i += 1
        with torch.no_grad():
            X = clamp(theta).requires_grad_(True)
        loss = -acquisition_function(X).sum()
        grad = torch.autograd.grad(loss, X)[0]
        if callback:
            callback(i, loss, grad, X)
        opt.zero_grad()
        theta.grad = grad
        opt.step()
        if decay:
            opt = optimizer(params=[theta], lr=lr / i ** 0.7)
        stop = stopping.evaluate(fvals=loss.detach())
    out = clamp(theta.detach())
    with torch.no_grad():
        vals = acquisition_function(out)
    return out, vals


def gen_candidates_scipy(
    initial_conditions: Tensor,
    acquisition_function,
    lower_bounds: Optional[Union[float, Tensor]] = None,
    upper_bounds: Optional[Union[float, Tensor]] = None,
    inequality_constraints=None,
    equality_constraints=None,
    nonlinear_inequality_constraints=None,
    options: Optional[Dict] = None,
    fixed_features: Optional[Dict[int, Optional[float]]] = None,
) -> Tuple[Tensor, Tensor]:
    """``gen.py:37-261`` restricted to box-bounded L-BFGS-B with exact gradients (what ``optimize_acqf_mixed`` needs for
    the ``enumerate__*`` style of use).  Values and gradients come from the device; scipy drives the iteration on the host.
    Constraints and the ``with_grad=False`` finite-difference mode are not part of the PR path (raise)."""
    from scipy.optimize import minimize

    if inequality_constraints or equality_constraints or nonlinear_inequality_constraints:
        raise NotImplementedError("parameter constraints are outside the PR hot path")
    options = dict(options or {})
    if not options.pop("with_grad", True):
        raise NotImplementedError("with_grad=False (finite differences) is a baseline mode, not part of the PR path")
    if fixed_features:
        d = initial_conditions.shape[-1]
        cols = sorted(fixed_features)
        free = [i for i in range(d) if i not in fixed_features]
        sub = FixedFeatureAcquisitionFunction(acquisition_function, d, cols, [fixed_features[c] for c in cols])
        sl = lambda b_: b_[..., free] if torch.is_tensor(b_) and b_.ndim > 0 else b_  # noqa: E731
        cand, vals = gen_candidates_scipy(initial_conditions[..., free], sub, sl(lower_bounds), sl(upper_bounds),
                                          options=options)
        return sub._construct_X_full(cand), vals
    shape = initial_conditions.shape
    x0 = columnwise_clamp(initial_conditions, lower_bounds, upper_bounds).detach()
    lo = None if lower_bounds is None else torch.as_tensor(lower_bounds, dtype=x0.dtype).cpu().expand(shape).reshape(-1)
    hi = None if upper_bounds is None else torch.as_tensor(upper_bounds, dtype=x0.dtype).cpu().expand(shape).reshape(-1)
    bnds = [(None if lo is None else float(lo[i]), None if hi is None else float(hi[i])) for i in range(x0.numel())]

    def f(x):
        if np.isnan(x).any():
            raise RuntimeError(f"{np.isnan(x).sum()} elements of the {x.size} element array `x` are NaN.")
        X = torch.from_numpy(x).to(x0).view(shape).contiguous().requires_grad_(True)
        loss = -acquisition_function(X).sum()
        g = torch.autograd.grad(loss, X)[0].reshape(-1).cpu().numpy().astype(np.float64)
        if np.isnan(g).any():
            raise RuntimeError(f"{np.isnan(g).sum()} elements of the {x.size} element gradient array are NaN.")
        return float(loss.detach()), g

    res = minimize(f, x0.reshape(-1).cpu().numpy().astype(np.float64), method=options.pop("method", "L-BFGS-B"),
                   jac=True, bounds=bnds, options={k: v for k, v in options.items() if k in ("maxiter", "ftol", "gtol",
                                                                                             "maxfun", "maxcor")})
    cand = columnwise_clamp(torch.from_numpy(res.x).to(x0).view(shape), lower_bounds, upper_bounds,
                            raise_on_violation=True)
    with torch.no_grad():
        vals = acquisition_function(cand)
    return cand, vals


def get_best_candidates(batch_candidates: Tensor, batch_values: Tensor) -> Tensor:
    """``gen.py:346-374``."""
    return batch_candidates[torch.argmax(batch_values.view(-1), dim=0)]
