"""Exact-GP model surface (``FixedNoiseGP``) and the device-side GP state the CUDA path reads.

The reference builds a BoTorch ``FixedNoiseGP`` with a ``ConstantMean`` and the ``get_kernel`` covariance
(``discrete_mixed_bo/experiment_utils.py:237-351``) and lets GPyTorch's exact prediction strategy cache
``(K + Sigma)^-1 (y - c)`` and the inverse Cholesky root.  ``GPState`` is that cache as plain device tensors
(``X_train``, ``Linv``, ``alpha``, constant, kernel structure) handed to ``libprb200`` through ``prb_model_create``.
The whole build -- ``K(X, X)``, blocked Cholesky, triangular inverse, ``alpha``, packing -- runs in the library
(``prb_model_create_from_data``, ``csrc/prb_chol.cu``: FP64 DMMA block GEMMs, no cuSOLVER / cuBLAS); the step *before* the
hot path.  ``builder="torch"`` keeps the library factorisation as a cross-check for the tests.
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import List, Optional, Sequence

import torch
from torch import Tensor

from . import _capi as capi
from .kernels import TermT, kernel_fingerprint, kernel_terms

MIN_FIXED_NOISE = 1e-6  # gpytorch raises fixed noise below this (float64) with a warning
MIN_VARIANCE = 1e-10  # gpytorch.settings.min_variance for float64


def _fill_model_desc(desc: capi.PrbModelDesc, terms: Sequence[TermT], n: int, d_k: int) -> None:
    if len(terms) > capi.PRB_MAX_TERMS:
        raise NotImplementedError(f"at most {capi.PRB_MAX_TERMS} additive terms are supported")
    desc.n, desc.d_k, desc.n_terms = n, d_k, len(terms)
    for t, (scale, additive, factors) in enumerate(terms):
        if len(factors) > capi.PRB_MAX_FACTORS:
            raise NotImplementedError(f"at most {capi.PRB_MAX_FACTORS} factors per term are supported")
        T = desc.terms[t]
        T.outputscale, T.additive, T.n_factors = float(scale), int(bool(additive)), len(factors)
        for f, (kind, dims, ls, power) in enumerate(factors):
            F = T.factors[f]
            F.kind = capi.PRB_FACTOR_MATERN52 if kind == "matern52" else capi.PRB_FACTOR_CATEGORICAL
            F.n_dims, F.power = len(dims), int(power)
            for i, (dd, l) in enumerate(zip(dims, ls)):
                F.dims[i], F.lengthscale[i] = int(dd), float(l)


class DeviceState:
    """A ``prb_model`` handle plus the per-model caches every entry point needs (workspace, options)."""

    def _init_common(self, device: torch.device, d_k: int) -> None:
        self.lib = capi.load_library()
        self.device = device
        self.d_k = int(d_k)
        self.handle = C.c_void_p()
        self._ws: Optional[Tensor] = None
        self._ws_need = {}
        self._ones: Optional[Tensor] = None
        self._options = {}
        self.chunk_cols = 0

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                self.lib.prb_model_destroy(h)
            except Exception:
                pass
            self.handle = None

    def set_option(self, option: int, value: int) -> None:
        if self._options.get(int(option)) == int(value):
            return
        self._options[int(option)] = int(value)
        capi.check(self.lib.prb_model_set_option(self.handle, int(option), int(value)), "prb_model_set_option")

    def ones(self, b: int) -> Tensor:
        """Cached all-ones upstream gradient (the library evaluates the estimator for grad_output = 1)."""
        t = self._ones
        if t is None or t.numel() < b:
            t = self._ones = torch.ones(max(b, 64), dtype=torch.float64, device=self.device)
        return t[:b]

    def workspace(self, n_points: int, d_theta: int, n_restarts: int = 0) -> Tensor:
        key = (int(n_points), int(n_restarts), int(d_theta), int(self.chunk_cols))
        need = self._ws_need.get(key)
        if need is None:
            need = self._ws_need[key] = int(self.lib.prb_workspace_bytes(self.handle, *key))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    # ---- base acquisition / posterior at explicit kernel-input points -------------------------------------------------
    def acq_points(self, acq_desc: capi.PrbAcqDesc, xk: Tensor, want_grad: bool = False, want_post: bool = False):
        capi.require_cuda(xk, "X")
        xk = xk.detach().to(torch.float64).contiguous()
        if xk.shape[-1] != self.d_k:
            raise ValueError(f"expected inputs with {self.d_k} columns, got {xk.shape[-1]}")
        B = xk.numel() // self.d_k
        out = torch.empty(B, dtype=torch.float64, device=self.device)
        mean = torch.empty(B, dtype=torch.float64, device=self.device) if want_post else None
        var = torch.empty(B, dtype=torch.float64, device=self.device) if want_post else None
        grad = torch.empty(B, self.d_k, dtype=torch.float64, device=self.device) if want_grad else None
        ws = self.workspace(B, self.d_k)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            capi.check(self.lib.prb_acq_points(self.handle, C.byref(acq_desc), xk.data_ptr(), B, out.data_ptr(),
                                               capi.ptr(mean), capi.ptr(var), capi.ptr(grad), ws.data_ptr(),
                                               ws.numel(), int(self.chunk_cols), st), "prb_acq_points")
        return out, mean, var, grad


class RFFState(DeviceState):
    """Device copy of one random-Fourier-feature posterior sample ``f(x) = scale * cos(x W + b) . w`` (``prb_rff_model_create``)."""

    def __init__(self, W: Tensor, bias: Tensor, weights: Tensor, scale: float):
        capi.require_cuda(W, "W")
        self._init_common(W.device, W.shape[0])
        self.W = W.detach().to(torch.float64).contiguous()
        self.bias = bias.detach().to(device=W.device, dtype=torch.float64).reshape(-1).contiguous()
        self.weights = weights.detach().to(device=W.device, dtype=torch.float64).reshape(-1).contiguous()
        self.scale = float(scale)
        m = self.W.shape[1]
        if self.bias.numel() != m or self.weights.numel() != m:
            raise ValueError("W [d, m], bias [m] and weights [m] do not agree")
        if self.d_k > capi.PRB_MAX_DIMS:
            raise NotImplementedError(f"input dimension {self.d_k} > {capi.PRB_MAX_DIMS}")
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            capi.check(self.lib.prb_rff_model_create(C.byref(self.handle), self.d_k, m, self.W.data_ptr(),
                                                     self.bias.data_ptr(), self.weights.data_ptr(), self.scale, stream),
                       "prb_rff_model_create")


class GPState(DeviceState):
    """Device-resident exact-GP caches + the ``prb_model`` handle.  One per (model, BO iteration)."""

    def __init__(self, train_X: Tensor, train_Y: Tensor, terms: Sequence[TermT], noise, mean_const: float,
                 builder: str = "device"):
        capi.require_cuda(train_X, "train_X")
        self._init_common(train_X.device, train_X.shape[-1])
        X = train_X.detach().to(torch.float64).contiguous()
        y = train_Y.detach().to(device=self.device, dtype=torch.float64).reshape(-1).contiguous()
        n, d_k = X.shape
        if d_k > capi.PRB_MAX_DIMS:
            raise NotImplementedError(f"kernel-input dimension {d_k} > {capi.PRB_MAX_DIMS}")
        self.n, self.d_k = n, d_k
        self.terms = list(terms)
        self.mean_const = float(mean_const)
        self.X_train, self.train_Y = X, y
        noise = torch.as_tensor(noise, dtype=torch.float64, device=self.device).reshape(-1)
        self.noise = noise.expand(n).contiguous() if noise.numel() == 1 else noise.contiguous()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            desc = capi.PrbModelDesc()
            _fill_model_desc(desc, self.terms, n, d_k)
            desc.X_train = X.data_ptr()
            desc.mean_const = self.mean_const
            handle = C.c_void_p()
            if builder == "device":
                # K(X, X), blocked Cholesky, triangular inverse, alpha and the packed factors: all in the library
                # (prb_chol.cu); Linv / alpha are exported for inspection and for the tests
                self.Linv = torch.empty(n, n, dtype=torch.float64, device=self.device)
                self.alpha = torch.empty(n, dtype=torch.float64, device=self.device)
                capi.check(self.lib.prb_model_create_from_data(C.byref(handle), C.byref(desc), y.data_ptr(),
                                                               self.noise.data_ptr(), self.Linv.data_ptr(),
                                                               self.alpha.data_ptr(), stream), "prb_model_create_from_data")
            elif builder == "torch":
                # cross-check path of the test-suite (cuSOLVER / cuBLAS factorisation, then prb_model_create packs it)
                K = torch.empty(n, n, dtype=torch.float64, device=self.device)
                capi.check(self.lib.prb_kernel_matrix(C.byref(desc), X.data_ptr(), n, K.data_ptr(), stream),
                           "prb_kernel_matrix")
                K.diagonal().add_(self.noise)
                L = torch.linalg.cholesky(K)
                self.Linv = torch.linalg.solve_triangular(
                    L, torch.eye(n, dtype=torch.float64, device=self.device), upper=False).contiguous()
                self.alpha = torch.cholesky_solve((y - self.mean_const).unsqueeze(-1), L).squeeze(-1).contiguous()
                desc.Linv, desc.alpha = self.Linv.data_ptr(), self.alpha.data_ptr()
                capi.check(self.lib.prb_model_create(C.byref(handle), C.byref(desc), stream), "prb_model_create")
            else:
                raise ValueError("builder must be 'device' or 'torch'")
        self.handle = handle
        self._desc = desc


class _ConstantMean:
    def __init__(self, constant: float = 0.0):
        self.constant = torch.tensor(float(constant), dtype=torch.float64)


class _FixedNoiseLikelihood:
    def __init__(self, noise: Tensor):
        self.noise = noise


class _Posterior:
    def __init__(self, mean: Tensor, variance: Tensor):
        self.mean, self.variance = mean, variance


class Model(torch.nn.Module):
    """Marker base class (``botorch.models.model.Model``)."""


class FixedNoiseGP(Model):
    """``botorch.models.FixedNoiseGP`` surface used by the reference (``experiment_utils.py:267-351``): ``train_inputs``,
    ``train_targets``, ``covar_module``, ``mean_module.constant``, ``likelihood.noise``, ``posterior(X)``.

    Hyperparameters are set by assignment on ``covar_module`` / ``mean_module`` (fitting the marginal likelihood is the
    reference's ``fit_gpytorch_model`` step, outside the hot path); call ``refresh()`` after changing them.
    ``input_transform`` may be a ``LatentCategoricalEmbedding`` (``kernel_type="mixed_latent"``,
    ``experiment_utils.py:289-313``): ``train_X`` and ``posterior(X)`` inputs are then in the numeric layout and the device
    state is built on the embedded inputs.  Other model-level transforms and multi-output models are not supported (raise)."""

    num_outputs = 1

    def __init__(self, train_X: Tensor, train_Y: Tensor, train_Yvar: Tensor, covar_module=None,
                 outcome_transform=None, input_transform=None, mean_const: float = 0.0):
        super().__init__()
        if outcome_transform is not None or (input_transform is not None and not hasattr(input_transform, "get_emb_table")):
            raise NotImplementedError("model-level transforms other than LatentCategoricalEmbedding are outside the B200 "
                                      "hot path")
        self.input_transform = input_transform
        if train_Y.ndim == 2 and train_Y.shape[-1] != 1:
            raise NotImplementedError("multi-output models are outside the B200 hot path")
        if covar_module is None:
            raise NotImplementedError("pass the covariance from get_kernel(); botorch_default is not implemented")
        capi.require_cuda(train_X, "train_X")
        self.train_inputs = (train_X.detach().to(torch.float64),)
        self.train_targets = train_Y.detach().to(torch.float64).reshape(-1)
        noise = train_Yvar.detach().to(torch.float64).reshape(-1)
        if bool((noise < MIN_FIXED_NOISE).any()):
            warnings.warn(f"Very small noise values detected; rounding up to {MIN_FIXED_NOISE} (GPyTorch behaviour).")
            noise = noise.clamp_min(MIN_FIXED_NOISE)
        self.likelihood = _FixedNoiseLikelihood(noise)
        self.covar_module = covar_module
        self.mean_module = _ConstantMean(mean_const)
        self._state: Optional[GPState] = None

    def refresh(self) -> "FixedNoiseGP":
        self._state = None
        return self

    def _fingerprint(self) -> tuple:
        return _model_fingerprint(self)

    @property
    def state(self) -> GPState:
        """The device GP state, rebuilt whenever the training data, the noise, the mean constant or any kernel
        hyper-parameter has changed since it was built (a stale Cholesky factor would silently give wrong PR values)."""
        fp = self._fingerprint()
        if self._state is None or self._state_fp != fp:
            X = self.train_inputs[0]
            if self.input_transform is not None:
                with torch.no_grad():
                    X = self.input_transform.to(X.device).transform(X).contiguous()
            self._state = GPState(X, self.train_targets, kernel_terms(self.covar_module, X.shape[-1]),
                                  self.likelihood.noise, float(self.mean_module.constant))
            self._state_fp = fp
        return self._state

    def posterior(self, X: Tensor, observation_noise: bool = False, **kwargs) -> _Posterior:
        if observation_noise:
            raise NotImplementedError
        desc = capi.PrbAcqDesc(kind=capi.PRB_ACQ_MEAN, ei_variant=0, param=0.0, min_variance=MIN_VARIANCE,
                               acq_var_floor=0.0)
        shape = X.shape[:-1]
        if self.input_transform is not None:
            X = self.input_transform.transform(X)
        _, mean, var, _ = self.state.acq_points(desc, X, want_post=True)
        return _Posterior(mean.reshape(*shape, 1), var.reshape(*shape, 1))


class ModelListGP(Model):
    """``botorch.models.ModelListGP`` surface the reference uses (``experiment_utils.py:337-346``): independent single-output
    GPs over the same inputs, one per outcome (objective + constraints).  ``models`` / ``num_outputs`` / ``posterior``."""

    def __init__(self, *gp_models):
        super().__init__()
        if len(gp_models) < 1 or len(gp_models) > capi.PRB_MAX_MODELS:
            raise NotImplementedError(f"1 .. {capi.PRB_MAX_MODELS} models are supported")
        self.models = torch.nn.ModuleList(gp_models)

    @property
    def num_outputs(self) -> int:
        return len(self.models)

    @property
    def train_inputs(self):
        return [m.train_inputs for m in self.models]

    def posterior(self, X: Tensor, observation_noise: bool = False, **kwargs) -> _Posterior:
        if observation_noise:
            raise NotImplementedError
        ps = [m.posterior(X) for m in self.models]
        return _Posterior(torch.cat([p.mean for p in ps], dim=-1), torch.cat([p.variance for p in ps], dim=-1))


class MultiState:
    """The device GP states of a model list plus the pointer array and the shared workspace the ``*_multi`` entry points take."""

    def __init__(self, states: Sequence[GPState]):
        self.states = list(states)
        s0 = self.states[0]
        if any(s.d_k != s0.d_k or s.n != s0.n or s.device != s0.device for s in self.states):
            raise NotImplementedError("the models of a list must share the training inputs' shape and the device")
        self.lib, self.device, self.d_k = s0.lib, s0.device, s0.d_k
        self.handles = (C.c_void_p * len(self.states))(*[s.handle for s in self.states])
        self._ws: Optional[Tensor] = None
        self._need = {}

    def __len__(self) -> int:
        return len(self.states)

    def workspace(self, n_points: int, d_theta: int, n_restarts: int = 0) -> Tensor:
        key = (int(n_points), int(n_restarts), int(d_theta))
        need = self._need.get(key)
        if need is None:
            need = self._need[key] = int(self.lib.prb_workspace_bytes_multi(self.handles, len(self.states), key[0], key[1],
                                                                            key[2]))
            if need == 0:
                raise capi.PrbError("prb_workspace_bytes_multi failed (model list not usable on the device path)")
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws


def multi_state_from_model(model) -> MultiState:
    """``MultiState`` of a ``ModelListGP`` (ours or duck-typed: anything with ``.models``); rebuilt when a member's state is."""
    members = list(model.models)
    if any(getattr(m, "input_transform", None) is not None for m in members):
        # every member would embed the sampled categories with ITS OWN tables; the model-list path shares one kernel input
        raise NotImplementedError("model lists whose members carry a LatentCategoricalEmbedding are not supported")
    states = [gp_state_from_model(m) for m in members]
    cached = getattr(model, "_prb_multi", None)
    if cached is not None and len(cached[0]) == len(states) and all(a is b for a, b in zip(cached[0], states)):
        return cached[1]
    ms = MultiState(states)
    try:
        model._prb_multi = (states, ms)
    except Exception:
        pass
    return ms


def _tensor_fp(t) -> tuple:
    if torch.is_tensor(t):
        return (t.data_ptr(), t._version, tuple(t.shape))
    return (t,)


def _model_fingerprint(model) -> tuple:
    """Identity of everything the device GP state is built from: storage + in-place version counter of the training
    tensors, the noise and the mean constant, and the kernel tree's hyper-parameters (``kernel_fingerprint``).  Re-assigning
    or mutating any of them in place changes it; a few microseconds per call."""
    mm = model.mean_module
    const = getattr(mm, "raw_constant", None)
    if const is None:
        const = getattr(mm, "constant", None)
    tf = getattr(model, "input_transform", None)
    return (_tensor_fp(model.train_inputs[0]), _tensor_fp(model.train_targets), _tensor_fp(model.likelihood.noise),
            _tensor_fp(const), kernel_fingerprint(model.covar_module),
            tf.fingerprint() if hasattr(tf, "fingerprint") else None)


def gp_state_from_model(model) -> GPState:
    """GP state of one of our models, or of a duck-typed BoTorch ``FixedNoiseGP``/``SingleTaskGP`` in eval mode
    (``train_inputs[0]``, ``train_targets``, ``mean_module.constant``, ``likelihood.noise``, ``covar_module`` tree)."""
    if isinstance(model, FixedNoiseGP) or isinstance(getattr(model, "state", None), DeviceState):
        return model.state
    for attr in ("input_transform", "outcome_transform"):
        if getattr(model, attr, None) is not None:
            raise NotImplementedError(f"model.{attr} is not supported by the B200 hot path")
    if hasattr(model, "models"):
        raise NotImplementedError("a ModelListGP has one device state per member: use multi_state_from_model")
    fp = _model_fingerprint(model)
    cached = getattr(model, "_prb_state", None)
    if cached is not None and cached[0] == fp:
        return cached[1]
    X = model.train_inputs[0]
    if X.ndim != 2:
        raise NotImplementedError("batched models are not supported by the B200 hot path")
    noise = model.likelihood.noise.detach().reshape(-1)
    state = GPState(X, model.train_targets, kernel_terms(model.covar_module, X.shape[-1]), noise,
                    float(torch.as_tensor(model.mean_module.constant).detach().reshape(-1)[0]))
    try:
        model._prb_state = (fp, state)
    except Exception:
        pass
    return state
