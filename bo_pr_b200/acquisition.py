"""Base acquisition functions evaluated by ``libprb200``.

Surface of ``botorch.acquisition.analytic.{ExpectedImprovement, UpperConfidenceBound, PosteriorMean}`` as the reference
uses it (``discrete_mixed_bo/experiment_utils.py:391-393, 476-480, 489-492``): ``acq(X[..., 1, d]) -> [...]`` with
``.model``, ``.best_f`` / ``.beta``.  ``forward`` is differentiable w.r.t. ``X`` (the gradient comes from the device
path: second triangular product + kernel-derivative contraction).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import _capi as capi
from .models import MIN_VARIANCE, GPState, gp_state_from_model, multi_state_from_model


class AcquisitionFunction(torch.nn.Module):
    """``botorch.acquisition.acquisition.AcquisitionFunction`` stand-in."""

    def __init__(self, model=None):
        super().__init__()
        if model is not None:
            self.model = model
        self.X_pending = None

    def set_X_pending(self, X_pending: Optional[Tensor] = None) -> None:
        self.X_pending = X_pending


class _PointsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X: Tensor, state: GPState, desc: capi.PrbAcqDesc):
        need = X.requires_grad
        val, _, _, grad = state.acq_points(desc, X.reshape(-1, X.shape[-1]), want_grad=need)
        ctx.grad = grad
        ctx.xshape = X.shape
        return val.reshape(X.shape[:-1])

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.grad is None:
            return None, None, None
        g = ctx.grad.reshape(ctx.xshape) * grad_output.unsqueeze(-1)
        return g, None, None


class AnalyticAcquisitionFunction(AcquisitionFunction):
    """q = 1 analytic AF on a single-output exact GP.  ``posterior_transform`` is not supported (raises)."""

    def __init__(self, model, posterior_transform=None, **kwargs):
        super().__init__(model=model)
        if posterior_transform is not None:
            raise NotImplementedError("posterior transforms are outside the B200 hot path")

    @property
    def state(self) -> GPState:
        return gp_state_from_model(self.model)

    def acq_desc(self) -> capi.PrbAcqDesc:  # pragma: no cover - abstract
        raise NotImplementedError

    def forward(self, X: Tensor) -> Tensor:
        if X.shape[-2] != 1:
            raise AssertionError(f"Expected X to be `batch_shape x q=1 x d`, but got X with shape {tuple(X.shape)}.")
        tf = getattr(self.model, "input_transform", None)
        if tf is not None:  # LatentCategoricalEmbedding: numeric layout -> kernel input (torch gather, differentiable in
            X = tf.transform(X)  # the non-categorical columns)
        return _PointsFn.apply(X.squeeze(-2), self.state, self.acq_desc())


class ExpectedImprovement(AnalyticAcquisitionFunction):
    """EI = sigma (phi(u) + u Phi(u)), u = (mu - best_f) / sigma, sigma = sqrt(max(var, var_floor)).
    ``variant="erf"`` / ``var_floor=1e-9`` is BoTorch <= 0.8 (the reference's era); ``"erfc"`` the later helper."""

    def __init__(self, model, best_f, posterior_transform=None, maximize: bool = True, var_floor: float = 1e-9,
                 variant: str = "erf", **kwargs):
        super().__init__(model=model, posterior_transform=posterior_transform)
        if not maximize:
            raise NotImplementedError("maximize=False is not used by the reference")
        self.maximize = maximize
        self.register_buffer("best_f", torch.as_tensor(best_f, dtype=torch.float64).detach().clone())
        self.var_floor = float(var_floor)
        self.variant = variant

    def acq_desc(self) -> capi.PrbAcqDesc:
        # float(best_f) is a device->host read when the buffer lives on the GPU: redo it only when best_f changes
        ver = (self.best_f.data_ptr(), self.best_f._version, self.variant, self.var_floor)
        if getattr(self, "_desc_ver", None) != ver:
            self._desc = capi.PrbAcqDesc(kind=capi.PRB_ACQ_EI, ei_variant=0 if self.variant == "erf" else 1,
                                         param=float(self.best_f), min_variance=MIN_VARIANCE,
                                         acq_var_floor=self.var_floor)
            self._desc_ver = ver
        return self._desc


class UpperConfidenceBound(AnalyticAcquisitionFunction):
    """UCB = mu + sqrt(beta var)  (``experiment_utils.py:476-480``; beta = 0.2 d ln(2 iter), ``:446-447``)."""

    def __init__(self, model, beta, posterior_transform=None, maximize: bool = True, **kwargs):
        super().__init__(model=model, posterior_transform=posterior_transform)
        if not maximize:
            raise NotImplementedError("maximize=False is not used by the reference")
        self.maximize = maximize
        self.register_buffer("beta", torch.as_tensor(beta, dtype=torch.float64).detach().clone())

    def acq_desc(self) -> capi.PrbAcqDesc:
        ver = (self.beta.data_ptr(), self.beta._version)
        if getattr(self, "_desc_ver", None) != ver:
            self._desc = capi.PrbAcqDesc(kind=capi.PRB_ACQ_UCB, ei_variant=0, param=float(self.beta),
                                         min_variance=MIN_VARIANCE, acq_var_floor=0.0)
            self._desc_ver = ver
        return self._desc


class PosteriorMean(AnalyticAcquisitionFunction):
    def __init__(self, model, posterior_transform=None, maximize: bool = True, **kwargs):
        super().__init__(model=model, posterior_transform=posterior_transform)
        self.maximize = maximize

    def acq_desc(self) -> capi.PrbAcqDesc:
        return capi.PrbAcqDesc(kind=capi.PRB_ACQ_MEAN, ei_variant=0, param=0.0, min_variance=MIN_VARIANCE,
                               acq_var_floor=0.0)


class ModelListAcquisitionFunction(AcquisitionFunction):
    """Acquisition functions over a ``ModelListGP`` (one independent single-output GP per outcome): evaluated by
    ``prb_acq_points_multi`` (values; the reference only differentiates them THROUGH a PR wrapper, which has its own entry
    point ``prb_mc_value_grad_multi``).  Subclasses provide ``multi_desc()``."""

    def __init__(self, model):
        super().__init__(model=model)
        if not hasattr(model, "models"):
            raise NotImplementedError(f"{type(self).__name__} needs a ModelListGP (one GP per outcome)")

    @property
    def state(self):
        return multi_state_from_model(self.model)

    def multi_desc(self) -> capi.PrbMultiAcqDesc:  # pragma: no cover - abstract
        raise NotImplementedError

    def forward(self, X: Tensor) -> Tensor:
        if X.shape[-2] != 1:
            raise AssertionError(f"Expected X to be `batch_shape x q=1 x d`, but got X with shape {tuple(X.shape)}.")
        if X.requires_grad:
            raise NotImplementedError(f"differentiate {type(self).__name__} through a PR wrapper")
        import ctypes as C
        ms = self.state
        capi.require_cuda(X, "X")
        xk = X.detach().to(torch.float64).reshape(-1, X.shape[-1]).contiguous()
        if xk.shape[-1] != ms.d_k:
            raise ValueError(f"expected inputs with {ms.d_k} columns, got {xk.shape[-1]}")
        n = xk.shape[0]
        out = torch.empty(n, dtype=torch.float64, device=X.device)
        ws = ms.workspace(n, ms.d_k, 1)
        desc = self.multi_desc()
        with torch.cuda.device(X.device):
            st = torch.cuda.current_stream().cuda_stream
            capi.check(ms.lib.prb_acq_points_multi(ms.handles, len(ms), C.byref(desc), xk.data_ptr(), n, out.data_ptr(), None,
                                                   None, ws.data_ptr(), ws.numel(), st), "prb_acq_points_multi")
        return out.reshape(X.shape[:-2])


class ConstrainedExpectedImprovement(ModelListAcquisitionFunction):
    """``botorch.acquisition.analytic.ConstrainedExpectedImprovement`` on a model list, as ``get_EI`` builds it for the
    constrained problems (``experiment_utils.py:354-392``): ``constraints = {output index: (lower or None, upper or None)}``."""

    def __init__(self, model, best_f, objective_index: int, constraints, maximize: bool = True, sigma_floor: float = 1e-9):
        super().__init__(model=model)
        if not maximize:
            raise NotImplementedError("maximize=False is not used by the reference")
        self.maximize = maximize
        self.objective_index = int(objective_index)
        self.constraints = {int(k): (None if v[0] is None else float(v[0]), None if v[1] is None else float(v[1]))
                            for k, v in dict(constraints).items()}
        if self.objective_index in self.constraints or any(k < 0 or k >= len(model.models) for k in self.constraints):
            raise ValueError("constraint indices must name non-objective outputs of the model list")
        self.register_buffer("best_f", torch.as_tensor(best_f, dtype=torch.float64).detach().clone())
        self.sigma_floor = float(sigma_floor)

    def multi_desc(self) -> capi.PrbMultiAcqDesc:
        ver = (self.best_f.data_ptr(), self.best_f._version)
        if getattr(self, "_desc_ver", None) != ver:
            d = capi.PrbMultiAcqDesc(kind=capi.PRB_MACQ_CEI, n_models=len(self.model.models),
                                     objective_index=self.objective_index, best_f=float(self.best_f),
                                     min_variance=MIN_VARIANCE, sigma_floor=self.sigma_floor)
            for k, (lo, hi) in self.constraints.items():
                d.has_lower[k], d.has_upper[k] = int(lo is not None), int(hi is not None)
                d.lower[k], d.upper[k] = (0.0 if lo is None else lo), (0.0 if hi is None else hi)
            self._desc, self._desc_ver = d, ver
        return self._desc


class ExpectedHypervolumeImprovement(ModelListAcquisitionFunction):
    """``botorch.acquisition.multi_objective.analytic.ExpectedHypervolumeImprovement`` (q = 1) as ``get_EHVI`` builds it
    (``experiment_utils.py:395-402``): every model of the list is an objective (maximised), ``partitioning`` a box
    decomposition of the non-dominated region above ``ref_point`` (``multi_objective.FastNondominatedPartitioning``; any
    object with ``get_hypercell_bounds() -> [2, n_cells, m]`` works).  The posterior variance is clamped at ``1e-9`` before
    the square root and infinite upper cell bounds at ``1e10``, as BoTorch does."""

    def __init__(self, model, ref_point, partitioning, objective=None, posterior_transform=None, var_floor: float = 1e-9):
        super().__init__(model=model)
        if objective is not None or posterior_transform is not None:
            raise NotImplementedError("objectives / posterior transforms are outside the B200 hot path")
        m = len(model.models)
        ref = torch.as_tensor(ref_point, dtype=torch.float64).reshape(-1)
        if ref.numel() != m:
            raise ValueError(f"ref_point has {ref.numel()} entries, the model list {m} outputs")
        if m < 2 or m > capi.PRB_MAX_MODELS:
            raise NotImplementedError(f"EHVI needs 2..{capi.PRB_MAX_MODELS} objectives")
        device = model.models[0].train_inputs[0].device
        cells = torch.as_tensor(partitioning.get_hypercell_bounds(), dtype=torch.float64)
        if cells.ndim != 3 or cells.shape[0] != 2 or cells.shape[-1] != m:
            raise ValueError("partitioning.get_hypercell_bounds() must be [2, n_cells, m]")
        self.partitioning = partitioning
        self.register_buffer("ref_point", ref.to(device))
        self.register_buffer("_cells", cells.to(device).contiguous())
        self.var_floor = float(var_floor)

    @property
    def cell_lower_bounds(self) -> Tensor:
        return self._cells[0]

    @property
    def cell_upper_bounds(self) -> Tensor:
        return self._cells[1]

    def multi_desc(self) -> capi.PrbMultiAcqDesc:
        ver = (self._cells.data_ptr(), self._cells._version)
        if getattr(self, "_desc_ver", None) != ver:
            self._desc = capi.PrbMultiAcqDesc(kind=capi.PRB_MACQ_EHVI, n_models=len(self.model.models), objective_index=0,
                                              best_f=0.0, min_variance=MIN_VARIANCE, sigma_floor=0.0,
                                              var_floor=self.var_floor, n_cells=int(self._cells.shape[1]),
                                              cell_bounds=self._cells.data_ptr())
            self._desc_ver = ver
        return self._desc


def acq_desc_of(acq_function) -> capi.PrbAcqDesc:
    """Descriptor for one of the classes above or a duck-typed BoTorch analytic AF (class name + ``best_f``/``beta``)."""
    if isinstance(acq_function, AnalyticAcquisitionFunction):
        return acq_function.acq_desc()
    if getattr(acq_function, "posterior_transform", None) is not None:
        raise NotImplementedError("posterior transforms are outside the B200 hot path")
    name = type(acq_function).__name__
    if name == "ExpectedImprovement":
        if not getattr(acq_function, "maximize", True):
            raise NotImplementedError
        return capi.PrbAcqDesc(kind=capi.PRB_ACQ_EI, ei_variant=0, param=float(acq_function.best_f),
                               min_variance=MIN_VARIANCE, acq_var_floor=1e-9)
    if name == "UpperConfidenceBound":
        return capi.PrbAcqDesc(kind=capi.PRB_ACQ_UCB, ei_variant=0, param=float(acq_function.beta),
                               min_variance=MIN_VARIANCE, acq_var_floor=0.0)
    if name == "PosteriorMean":
        return capi.PrbAcqDesc(kind=capi.PRB_ACQ_MEAN, ei_variant=0, param=0.0, min_variance=MIN_VARIANCE,
                               acq_var_floor=0.0)
    raise NotImplementedError(f"base acquisition function {name} has no device implementation in libprb200")


def is_nonnegative(acq_function) -> bool:
    """``botorch.optim.initializers.is_nonnegative`` for the AFs above (used at ``run_one_replication.py:407-408``)."""
    return type(acq_function).__name__ in ("ExpectedImprovement", "ConstrainedExpectedImprovement",
                                           "ExpectedHypervolumeImprovement")
