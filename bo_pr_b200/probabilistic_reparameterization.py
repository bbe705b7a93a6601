"""Probabilistic-reparameterization acquisition wrappers on the B200 path.

Drop-in for ``discrete_mixed_bo/probabilistic_reparameterization.py``: same class names, constructor signatures,
attributes (``acq_function``, ``model``, ``X_baseline``, ``input_transform``, ``one_hot_to_numeric``, ``ma_counter``,
``ma_hidden``, ``sample_candidates``) and the same 12-argument ``torch.autograd.Function``.  ``forward`` is one call into
``libprb200`` (``prb_mc_value_grad`` / ``prb_analytic_value_grad``): sampler -> cross-covariance panel -> FP64 DMMA
triangular products -> acquisition epilogue -> score-function / pathwise gradient reduction.  There is no torch
re-implementation of that arithmetic here and no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
from torch import Tensor
from torch.autograd import Function
from torch.nn.functional import one_hot

from . import _capi as capi
from .acquisition import AcquisitionFunction, ModelListAcquisitionFunction, acq_desc_of
from .input import (
    AnalyticProbabilisticReparameterizationInputTransform,
    ChainedInputTransform,
    MCProbabilisticReparameterizationInputTransform,
    Normalize,
    OneHotToNumeric,
    make_space_desc,
)
from .models import gp_state_from_model
from .wrapper import AcquisitionFunctionWrapper


def get_probabilistic_reparameterization_input_transform(
    dim: int,
    integer_indices: List[int],
    integer_bounds: Tensor,
    categorical_features: Optional[Dict[int, int]] = None,
    use_analytic: bool = False,
    mc_samples: int = 1024,
    resample: bool = False,
    flip: bool = False,
    tau: float = 0.1,
    rng: str = "sobol",
    philox_seed: int = 0,
) -> ChainedInputTransform:
    """``probabilistic_reparameterization.py:36-96``: ``{unnormalize -> round -> normalize}`` in ``eval()`` mode; the
    Normalize pair acts on the integer columns only and is omitted when there are none."""
    bounds = torch.zeros(2, dim, dtype=integer_bounds.dtype, device=integer_bounds.device)
    bounds[1] = 1
    bounds[:, integer_indices] = integer_bounds
    tfs = OrderedDict()
    has_ints = integer_indices is not None and len(integer_indices) > 0
    if has_ints:
        tfs["unnormalize"] = Normalize(d=dim, bounds=bounds, indices=integer_indices, transform_on_train=False,
                                       transform_on_eval=True, transform_on_fantasize=False, reverse=True)
    if use_analytic:
        tfs["round"] = AnalyticProbabilisticReparameterizationInputTransform(
            dim=dim, integer_indices=integer_indices, integer_bounds=integer_bounds,
            categorical_features=categorical_features, tau=tau)
    else:
        tfs["round"] = MCProbabilisticReparameterizationInputTransform(
            integer_indices=integer_indices, integer_bounds=integer_bounds, categorical_features=categorical_features,
            resample=resample, mc_samples=mc_samples, flip=flip, tau=tau, rng=rng, philox_seed=philox_seed)
    if has_ints:
        tfs["normalize"] = Normalize(d=dim, bounds=bounds, indices=integer_indices, transform_on_train=False,
                                     transform_on_eval=True, transform_on_fantasize=False, reverse=False)
    tf = ChainedInputTransform(**tfs)
    tf.eval()
    return tf


class AbstractProbabilisticReparameterization(AcquisitionFunctionWrapper, ABC):
    """``probabilistic_reparameterization.py:99-238``."""

    def __init__(
        self,
        acq_function: AcquisitionFunction,
        dim: int,
        integer_indices: Optional[List[int]] = None,
        integer_bounds: Optional[Tensor] = None,
        categorical_features: Optional[Dict[int, int]] = None,
        batch_limit: Optional[int] = None,
        apply_numeric: bool = False,
        **kwargs,
    ) -> None:
        if categorical_features is None and (integer_indices is None or integer_bounds is None):
            raise NotImplementedError("categorical_features or integer indices and integer_bounds must be provided.")
        super().__init__(acq_function=acq_function)
        self.batch_limit = batch_limit  # accepted for API parity; the fused path never materialises b x mc x n
        self.dim = dim
        self.apply_numeric = bool(apply_numeric)
        if integer_bounds is None:
            raise NotImplementedError("integer_bounds must be a tensor (pass a [2, 0] tensor when there are no "
                                      "integers; the reference dereferences integer_bounds.device too)")
        device = integer_bounds.device
        if apply_numeric:
            self.one_hot_to_numeric = OneHotToNumeric(categorical_features=categorical_features,
                                                      transform_on_train=False, transform_on_eval=True,
                                                      transform_on_fantasize=False)
            self.one_hot_to_numeric.eval()
        else:
            self.one_hot_to_numeric = None
        discrete_indices: List[int] = []
        ints = list(integer_indices) if integer_indices is not None else []
        self.register_buffer("integer_indices", torch.tensor(ints, dtype=torch.long, device=device))
        self.register_buffer("integer_bounds", integer_bounds if integer_indices is not None
                             else torch.tensor([], dtype=integer_bounds.dtype, device=device))
        discrete_indices.extend(ints)
        if categorical_features is not None and len(categorical_features) > 0:
            categorical_indices = list(range(min(categorical_features.keys()), dim))
            discrete_indices.extend(categorical_indices)
            self.categorical_features = categorical_features
        else:
            categorical_indices = []
            self.categorical_features = None
        self.register_buffer("categorical_indices", torch.tensor(categorical_indices, dtype=torch.long, device=device))
        self.register_buffer("cont_indices", torch.tensor(sorted(set(range(dim)) - set(discrete_indices)),
                                                          dtype=torch.long, device=device))
        # moving-average baseline state: (ma_counter, ma_hidden) live in ONE device buffer so the library can read the
        # pre-call state and advance it in place without a host round trip (:185-196, :499-505)
        self.register_buffer("_ma_state", torch.zeros(2, dtype=torch.double, device=device))
        self.register_buffer("ma_baseline", torch.zeros(1, dtype=torch.double, device=device))
        self._tau = kwargs.get("tau", 0.1)

    @property
    def ma_counter(self) -> Tensor:
        return self._ma_state[0:1]

    @property
    def ma_hidden(self) -> Tensor:
        return self._ma_state[1:2]

    def _space(self) -> capi.PrbSpaceDesc:
        tau = self.input_transform["round"].tau
        # a model-level LatentCategoricalEmbedding (kernel_type "mixed_latent") is applied by the sampler kernel
        latent = getattr(getattr(self.acq_function, "model", None), "input_transform", None)
        if latent is not None and not hasattr(latent, "device_tables"):
            latent = None
        key = (tau, None if latent is None else latent.fingerprint())
        cached = self.__dict__.get("_space_cached")
        if cached is None or cached[0] != key:
            sp = make_space_desc(self.dim, self.integer_indices.tolist(),
                                 self.integer_bounds if self.integer_indices.numel() > 0 else None,
                                 self.categorical_features, tau, apply_numeric=self.apply_numeric, latent=latent,
                                 device=self._ma_state.device)
            if latent is not None:
                key = (tau, latent.fingerprint())
            cached = self.__dict__["_space_cached"] = (key, sp)
        return cached[1]

    def sample_candidates(self, X: Tensor, rng: str = "torch", seed: Optional[int] = None, offset: int = 1 << 32,
                          draw_device=None, u: Optional[Tensor] = None) -> Tensor:
        """``:198-233``: one discrete draw per restart from p(theta*).  ``X [b, 1, d]`` -> ``[b, 1, d]``.

        ``rng="torch"`` (default, the reference's behaviour): the rounding probabilities come from the device
        (``prb_pr_sample``), the draws from torch's global generator in the reference's own order -- one
        ``Bernoulli(probs=p).sample()`` per integer parameter (:204-209), then one ``Categorical(probs=p).sample()`` per
        one-hot block (:219-229) -- on ``draw_device`` (default: ``X.device``, i.e. the CUDA generator, as the reference
        running on the same GPU; ``"cpu"`` reproduces a CPU run of the reference bit for bit, which is what the golden
        fixtures record).  ``rng="philox"``: the whole draw is ONE kernel (``prb_sample_candidates``) on the counter-based
        stream ``(seed, offset)``, or on explicit uniforms ``u [b, n_int + n_cat]``; no host round trip."""
        _check_X(X, self.dim)
        if rng == "philox" or u is not None:
            return self._sample_candidates_device(X, 0 if seed is None else int(seed), int(offset), u)[0]
        if rng != "torch":
            raise ValueError("rng must be 'torch' or 'philox'")
        tf = self.input_transform
        x_un = tf["unnormalize"](X) if "unnormalize" in tf else X.clone()
        prob = tf["round"].get_rounding_prob(X=x_un)
        dev = X.device if draw_device is None else torch.device(draw_device)
        x_un, prob = x_un.to(dev), prob.to(dev)
        ints = self.integer_indices.to(dev)
        col = 0
        for i in ints.tolist():  # one Bernoulli call per integer parameter: the generator is consumed in this order
            z = torch.distributions.Bernoulli(probs=prob[..., col]).sample()
            x_un[..., i] = x_un[..., i].floor() + z
            col += 1
        if ints.numel() > 0:
            ib = self.integer_bounds.to(dev)
            x_un[..., ints] = torch.minimum(torch.maximum(x_un[..., ints], ib[0]), ib[1])
        start = self.cont_indices.shape[0] + col  # first one-hot column
        if self.categorical_indices.shape[0] > 0:
            for cardinality in self.categorical_features.values():
                cat = torch.distributions.Categorical(probs=prob[..., col:col + cardinality]).sample()
                x_un[..., start:start + cardinality] = one_hot(cat, num_classes=cardinality).to(x_un)
                col += cardinality
                start += cardinality
        x_un = x_un.to(X.device)
        return tf["normalize"](x_un) if "normalize" in tf else x_un

    def _sample_candidates_device(self, X: Tensor, seed: int, offset: int, u: Optional[Tensor] = None,
                                  want_xk: bool = False):
        b, d = X.shape[0], X.shape[-1]
        theta = X.detach().to(torch.float64).reshape(b, d).contiguous()
        out = torch.empty(b, d, dtype=torch.float64, device=X.device)
        sp = self._space()
        xk = torch.empty(b, sp_dk(sp), dtype=torch.float64, device=X.device) if want_xk else None
        if u is not None:
            capi.require_cuda(u, "u")
            u = u.detach().to(torch.float64).reshape(b, -1).contiguous()
            if u.shape[1] != sp.n_int + sp.n_cat:
                raise ValueError(f"u must have {sp.n_int + sp.n_cat} columns (one per integer / categorical parameter)")
        lib = capi.load_library()
        with torch.cuda.device(X.device):
            st = torch.cuda.current_stream().cuda_stream
            capi.check(lib.prb_sample_candidates(C.byref(sp), theta.data_ptr(), b, capi.ptr(u), seed, offset,
                                                 out.data_ptr(), capi.ptr(xk), st), "prb_sample_candidates")
        return out.reshape(b, 1, d).to(X.dtype), xk

    @abstractmethod
    def forward(self, X: Tensor) -> Tensor:
        """Compute PR."""


# ----------------------------------------------------------------------------------------------------------------------
def sp_dk(sp: capi.PrbSpaceDesc) -> int:
    """Kernel-input width of a PR space: numeric layout when the base AF sees OneHotToNumeric inputs."""
    if sp.apply_numeric and sp.n_cat > 0:
        if sp.latent_tables:
            return sp.cat_start[0] + sum(sp.latent_dim[j] for j in range(sp.n_cat))
        return sp.cat_start[0] + sp.n_cat
    return sp.d


def _check_X(X: Tensor, dim: int):
    capi.require_cuda(X, "X")
    if X.ndim != 3 or X.shape[-2] != 1:
        raise NotImplementedError(f"expected X of shape b x q=1 x d, got {tuple(X.shape)} (q > 1 is not supported)")
    if X.shape[-1] != dim:
        raise ValueError(f"expected {dim} columns, got {X.shape[-1]}")


class _AnalyticPRFunction(Function):
    @staticmethod
    def forward(ctx, X: Tensor, module: "AnalyticProbabilisticReparameterization"):
        state = gp_state_from_model(module.acq_function.model)
        lib = state.lib
        b, d = X.shape[0], X.shape[-1]
        if b == 0:
            ctx.unit_grad = None
            return X.new_empty(0, dtype=torch.float64)
        rnd = module.input_transform["round"]
        codes = rnd.option_codes
        if codes.device != X.device:  # the transform lives on the constructor's `device`: never hand the kernel a foreign pointer
            codes = module.__dict__.get("_codes_on")
            if codes is None or codes.device != X.device:
                codes = module.__dict__["_codes_on"] = rnd.option_codes.to(X.device)
        n_opt = codes.shape[0]
        theta = X.detach().to(torch.float64).reshape(b, d).contiguous()
        need = ctx.needs_input_grad[0]
        value = torch.empty(b, dtype=torch.float64, device=X.device)
        grad = torch.empty(b, d, dtype=torch.float64, device=X.device) if need else None
        ones = torch.ones(b, dtype=torch.float64, device=X.device) if need else None
        ws = state.workspace(b * n_opt, d, b)
        sp, aq = module._space(), acq_desc_of(module.acq_function)
        with torch.cuda.device(X.device):
            st = torch.cuda.current_stream().cuda_stream
            capi.check(lib.prb_analytic_value_grad(state.handle, C.byref(sp), C.byref(aq), theta.data_ptr(), b,
                                                   codes.data_ptr(), n_opt, capi.ptr(ones), value.data_ptr(),
                                                   capi.ptr(grad), None, None, ws.data_ptr(), ws.numel(),
                                                   int(state.chunk_cols), st), "prb_analytic_value_grad")
        ctx.unit_grad = grad
        return value

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.unit_grad is None:
            return None, None
        return (ctx.unit_grad * grad_output.unsqueeze(-1)).unsqueeze(-2), None


class AnalyticProbabilisticReparameterization(AbstractProbabilisticReparameterization):
    """``probabilistic_reparameterization.py:241-320``: sum over all discrete configurations weighted by P(z | theta)."""

    def __init__(
        self,
        acq_function: AcquisitionFunction,
        dim: int,
        dtype: torch.dtype,
        device: torch.device,
        integer_indices: Optional[List[int]] = None,
        integer_bounds: Optional[Tensor] = None,
        categorical_features: Optional[Dict[int, int]] = None,
        batch_limit: Optional[int] = None,
        apply_numeric: bool = False,
        tau: float = 0.1,
    ) -> None:
        super().__init__(acq_function=acq_function, dim=dim, integer_indices=integer_indices,
                         integer_bounds=integer_bounds, categorical_features=categorical_features,
                         batch_limit=batch_limit, apply_numeric=apply_numeric)
        if dtype != torch.float64:
            raise NotImplementedError("libprb200 computes in float64 only (the reference runs in float64)")
        self.input_transform = get_probabilistic_reparameterization_input_transform(
            dim=dim, use_analytic=True, integer_indices=integer_indices if integer_indices is not None else [],
            integer_bounds=integer_bounds, categorical_features=categorical_features, tau=tau)
        self.input_transform.to(device=device)
        if self.batch_limit is None:
            self.batch_limit = self.input_transform["round"].all_discrete_options.shape[0]

    def forward(self, X: Tensor) -> Tensor:
        _check_X(X, self.dim)
        return _AnalyticPRFunction.apply(X, self)


class MCProbabilisticReparameterization(AbstractProbabilisticReparameterization):
    """``probabilistic_reparameterization.py:323-396``.  Extra keyword ``rng``: ``"sobol"`` (reference behaviour: cached
    scrambled-Sobol base samples) or ``"philox"`` (device-generated counter-based stream, ``philox_seed``)."""

    def __init__(
        self,
        acq_function: AcquisitionFunction,
        dim: int,
        integer_indices: Optional[List[int]] = None,
        integer_bounds: Optional[Tensor] = None,
        categorical_features: Optional[Dict[int, int]] = None,
        batch_limit: Optional[int] = None,
        apply_numeric: bool = False,
        mc_samples: int = 1024,
        grad_estimator: str = "reinforce",
        tau: float = 0.1,
        rng: str = "sobol",
        philox_seed: int = 0,
        dedup: bool = False,
    ) -> None:
        super().__init__(acq_function=acq_function, dim=dim, integer_indices=integer_indices,
                         integer_bounds=integer_bounds, categorical_features=categorical_features,
                         batch_limit=batch_limit, apply_numeric=apply_numeric)
        self.dedup = bool(dedup)  # PRB_OPT_DEDUP: evaluate each unique discrete configuration of a restart once
        if self.batch_limit is None:
            self.batch_limit = mc_samples
        if grad_estimator in ("arm", "u2g"):
            # the reference raises NameError here (undefined get_mc_probabilistic_reparameterization_input_transform,
            # :362-372); those estimators are dead code
            raise NotImplementedError(f"grad_estimator={grad_estimator!r} is dead code in the reference")
        if grad_estimator not in ("reinforce", "reinforce_ma"):
            raise ValueError(f"unknown grad_estimator {grad_estimator!r}")
        self.grad_estimator = grad_estimator
        self.mc_samples = mc_samples
        self._pr_acq_function = _MCProbabilisticReparameterization  # the autograd Function class (static methods)
        self.input_transform = get_probabilistic_reparameterization_input_transform(
            dim=dim, integer_indices=integer_indices if integer_indices is not None else [],
            integer_bounds=integer_bounds, categorical_features=categorical_features, mc_samples=mc_samples, tau=tau,
            rng=rng, philox_seed=philox_seed)
        self.input_transform_flip = None

    def screen(self, X: Tensor, chunk_rows: Optional[int] = None) -> Tensor:
        """Value-only evaluation of ``X [n, 1, d]`` in consecutive chunks of ``chunk_rows`` rows, one ``prb_mc_screen`` call
        (no host work between chunks).  Equivalent to ``torch.cat([self(X[s:s + chunk_rows]) for s in ...])`` under
        ``torch.no_grad()``, including the EMA-baseline state (one update per chunk)."""
        _check_X(X, self.dim)
        n, d = X.shape[0], X.shape[-1]
        chunk = n if chunk_rows is None else int(chunk_rows)
        out = torch.empty(n, dtype=torch.float64, device=X.device)
        if n == 0:
            return out
        state = gp_state_from_model(self.acq_function.model)
        state.set_option(capi.PRB_OPT_DEDUP, int(self.dedup))
        rnd = self.input_transform["round"]
        rnd.ensure_base_samples(1, X.device)
        u_int, u_cat = rnd.base_sample_ptrs()
        theta = X.detach().to(torch.float64).reshape(n, d).contiguous()
        rows = min(chunk, n)
        ws = state.workspace(rows * rnd.mc_samples, d, rows)
        est = capi.PRB_EST_REINFORCE_MA if self.grad_estimator == "reinforce_ma" else capi.PRB_EST_REINFORCE
        sp, aq = self._space(), acq_desc_of(self.acq_function)
        with torch.cuda.device(X.device):
            st = torch.cuda.current_stream().cuda_stream
            capi.check(state.lib.prb_mc_screen(state.handle, C.byref(sp), C.byref(aq), theta.data_ptr(), n, chunk,
                                               rnd.mc_samples, capi.ptr(u_int), capi.ptr(u_cat), 0, 0, est,
                                               self.ma_counter.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                               int(state.chunk_cols), st), "prb_mc_screen")
        return out

    def forward(self, X: Tensor) -> Tensor:
        """One ``prb_mc_value_grad`` enqueue.  The reference's ``forward`` hands twelve arguments to its autograd Function
        (:383-396; that Function is kept below with the same signature); this one passes the module itself, and
        everything that only depends on constructor arguments is resolved once per (module, device) -- the host time before
        the first kernel is queued is on the critical path of a sub-millisecond step (8 restarts per GPU)."""
        _check_X(X, self.dim)
        return _MCPRCall.apply(X, self)

    def value_and_grad_host(self, theta_host: Tensor, value_host: Tensor, grad_host: Optional[Tensor] = None,
                            device=None, synchronize: bool = True) -> None:
        """Host-buffer entry (``prb_mc_value_grad_host``): ``theta_host [b, 1, d]`` (float64, contiguous, ideally pinned) in,
        ``value_host [b]`` and ``grad_host [b, 1, d]`` (d value.sum() / d theta) out -- host->device copy, the fused step and
        the device->host copies in ONE library call on the current stream of ``device`` (default: the model's).  Same
        arithmetic and EMA-baseline side effects as ``forward`` + ``autograd.grad``; this is the path a caller without
        device tensors uses, and what ``bench.py`` times as ``e2e``."""
        state = gp_state_from_model(self.acq_function.model)
        dev = state.device if device is None else torch.device(device)
        for t, nm in ((theta_host, "theta_host"), (value_host, "value_host"), (grad_host, "grad_host")):
            if t is not None and (t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous()):
                raise ValueError(f"{nm} must be a contiguous float64 HOST tensor")
        b, d = theta_host.shape[0], theta_host.shape[-1]
        if theta_host.numel() != b * d or d != self.dim:
            raise ValueError(f"expected theta_host of shape b x 1 x {self.dim}")
        if b == 0:
            return
        cc = self._call_ctx(dev)
        rnd = cc["rnd"]
        state.set_option(capi.PRB_OPT_DEDUP, int(self.dedup))
        if rnd.resample or cc.get("u_ver") is None:
            rnd.ensure_base_samples(1, dev)
            ui, uc = rnd.base_sample_ptrs()
            cc["u_ver"] = (capi.ptr(ui), capi.ptr(uc), ui, uc)
        mc = rnd.mc_samples
        ws = state.workspace(b * mc, d, b)
        aq = acq_desc_of(self.acq_function)
        idx = cc["idx"]
        fn = capi.load_library().prb_mc_value_grad_host
        args = (state.handle, cc["sp"], C.byref(aq), theta_host.data_ptr(), b, mc, cc["u_ver"][0], cc["u_ver"][1], 0, 0,
                cc["est"], self._ma_state.data_ptr(), 1, int(grad_host is not None), value_host.data_ptr(),
                capi.ptr(grad_host), ws.data_ptr(), ws.numel(), int(synchronize))
        if torch._C._cuda_getDevice() == idx:
            capi.check(fn(*args, torch._C._cuda_getCurrentRawStream(idx)), "prb_mc_value_grad_host")
        else:
            with torch.cuda.device(dev):
                capi.check(fn(*args, torch._C._cuda_getCurrentRawStream(idx)), "prb_mc_value_grad_host")

    def _call_ctx(self, dev: torch.device):
        cc = self.__dict__.get("_cc")
        if cc is None or cc["dev"] != dev:
            rnd = self.input_transform["round"]
            cc = self.__dict__["_cc"] = {
                "dev": dev, "idx": dev.index if dev.index is not None else torch.cuda.current_device(), "rnd": rnd,
                "est": capi.PRB_EST_REINFORCE_MA if self.grad_estimator == "reinforce_ma" else capi.PRB_EST_REINFORCE,
                "sp": C.byref(self._space()), "sp_obj": self._space(), "tau": rnd.tau, "ma_ptr": self._ma_state.data_ptr(),
                "fn": capi.load_library().prb_mc_value_grad,
            }
        return cc

    def _launch(self, X: Tensor, need: bool):
        """(value [b], unit gradient [b, d] or None): the library call behind ``forward``."""
        b, d = X.shape[0], X.shape[-1]
        dev = X.device
        if b == 0:  # nothing to evaluate; the reference's EMA update would average an empty tensor (NaN): leave the state
            return X.new_empty(0, dtype=torch.float64), None
        if isinstance(self.acq_function, ModelListAcquisitionFunction):
            return self._launch_multi(X, need)
        cc = self._call_ctx(dev)
        rnd = cc["rnd"]
        if cc["tau"] != rnd.tau or cc["ma_ptr"] != self._ma_state.data_ptr():  # tau edited / module moved: rebuild
            self.__dict__.pop("_cc", None)
            self.__dict__.pop("_space_cached", None)
            cc = self._call_ctx(dev)
        state = gp_state_from_model(self.acq_function.model)
        state.set_option(capi.PRB_OPT_DEDUP, int(self.dedup))
        if rnd.resample or cc.get("u_ver") is None:
            rnd.ensure_base_samples(1, dev)
            ui, uc = rnd.base_sample_ptrs()
            cc["u_ver"] = (capi.ptr(ui), capi.ptr(uc), ui, uc)
        u_int, u_cat = cc["u_ver"][0], cc["u_ver"][1]
        theta = X if (X.dtype is torch.float64 and X.is_contiguous()) else \
            X.detach().to(torch.float64).reshape(b, d).contiguous()
        value = torch.empty(b, dtype=torch.float64, device=dev)
        grad = torch.empty(b, d, dtype=torch.float64, device=dev) if need else None
        ones = state.ones(b) if need else None
        mc = rnd.mc_samples
        ws = state.workspace(b * mc, d, b)
        aq = acq_desc_of(self.acq_function)
        idx = cc["idx"]
        args = (state.handle, cc["sp"], C.byref(aq), theta.data_ptr(), b, mc, u_int, u_cat, 0, 0, cc["est"],
                self._ma_state.data_ptr(), 1, capi.ptr(ones), value.data_ptr(), capi.ptr(grad), None, ws.data_ptr(),
                ws.numel(), int(state.chunk_cols))
        if torch._C._cuda_getDevice() == idx:
            capi.check(cc["fn"](*args, torch._C._cuda_getCurrentRawStream(idx)), "prb_mc_value_grad")
        else:
            with torch.cuda.device(dev):
                capi.check(cc["fn"](*args, torch._C._cuda_getCurrentRawStream(idx)), "prb_mc_value_grad")
        return value, grad


def _launch_multi(self, X: Tensor, need: bool):
    """Model-list acquisition functions (``ConstrainedExpectedImprovement``, ``ExpectedHypervolumeImprovement``):
    ``prb_mc_value_grad_multi``."""
    b, d = X.shape[0], X.shape[-1]
    dev = X.device
    acq = self.acq_function
    ms = acq.state
    rnd = self.input_transform["round"]
    rnd.ensure_base_samples(1, dev)
    u_int, u_cat = rnd.base_sample_ptrs()
    theta = X.detach().to(torch.float64).reshape(b, d).contiguous()
    value = torch.empty(b, dtype=torch.float64, device=dev)
    grad = torch.empty(b, d, dtype=torch.float64, device=dev) if need else None
    ones = torch.ones(b, dtype=torch.float64, device=dev) if need else None
    mc = rnd.mc_samples
    ws = ms.workspace(b * mc, d, b)
    est = capi.PRB_EST_REINFORCE_MA if self.grad_estimator == "reinforce_ma" else capi.PRB_EST_REINFORCE
    sp, desc = self._space(), acq.multi_desc()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        capi.check(ms.lib.prb_mc_value_grad_multi(ms.handles, len(ms), C.byref(sp), C.byref(desc), theta.data_ptr(), b, mc,
                                                  capi.ptr(u_int), capi.ptr(u_cat), 0, 0, est, self._ma_state.data_ptr(), 1,
                                                  capi.ptr(ones), value.data_ptr(), capi.ptr(grad), None, ws.data_ptr(),
                                                  ws.numel(), st), "prb_mc_value_grad_multi")
    return value, grad


MCProbabilisticReparameterization._launch_multi = _launch_multi


class _MCPRCall(Function):
    """Lean autograd bridge used by ``MCProbabilisticReparameterization.forward`` (two arguments instead of twelve)."""

    @staticmethod
    def forward(ctx, X: Tensor, module: "MCProbabilisticReparameterization"):
        value, ctx.unit_grad = module._launch(X, ctx.needs_input_grad[0])
        return value

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.unit_grad is None:
            return None, None
        # the library evaluated the estimator for grad_output = 1; it is linear in grad_output (:613-633)
        return (ctx.unit_grad * grad_output.unsqueeze(-1)).unsqueeze(-2), None


class _MCProbabilisticReparameterization(Function):
    """Same call signature as the reference's Function (``probabilistic_reparameterization.py:399-648``).

    ``ma_counter`` / ``ma_hidden`` must be the two adjacent one-element views of a single contiguous double[2] buffer
    (that is how ``AbstractProbabilisticReparameterization`` stores them): the library reads the pre-call state for the
    baseline and advances it in place."""

    @staticmethod
    def forward(ctx, input: Tensor, acq_function, input_tf, input_tf_flip, batch_limit, integer_indices, cont_indices,
                categorical_indices, grad_estimator, one_hot_to_numeric, ma_counter, ma_hidden):
        X = input
        rnd = input_tf["round"]
        d = X.shape[-1]
        _check_X(X, d)
        if grad_estimator not in ("reinforce", "reinforce_ma"):
            raise NotImplementedError(grad_estimator)
        if ma_hidden.data_ptr() != ma_counter.data_ptr() + 8:
            raise ValueError("ma_counter and ma_hidden must be adjacent views of one double[2] device buffer")
        state = gp_state_from_model(acq_function.model)
        lib = state.lib
        b = X.shape[0]
        if b == 0:  # nothing to evaluate; the reference's EMA update would average an empty tensor (NaN): leave the state
            ctx.unit_grad = None
            return X.new_empty(0, dtype=torch.float64)
        theta = X.detach().to(torch.float64).reshape(b, d).contiguous()
        rnd.ensure_base_samples(1, X.device)
        u_int, u_cat = rnd.base_sample_ptrs()
        mc = rnd.mc_samples
        # the space descriptor only depends on constructor arguments: build it once per transform (reading the index /
        # bounds buffers back from the device costs a host sync per call otherwise)
        key = (d, one_hot_to_numeric is not None, rnd.tau)
        cache = rnd.__dict__.setdefault("_space_desc_cache", {})
        sp = cache.get(key)
        if sp is None:
            ints = integer_indices.tolist()
            ib = rnd.integer_bounds if len(ints) > 0 else None
            sp = cache[key] = make_space_desc(d, ints, ib, rnd.categorical_features, rnd.tau,
                                              apply_numeric=one_hot_to_numeric is not None)
        aq = acq_desc_of(acq_function)
        need = ctx.needs_input_grad[0]
        value = torch.empty(b, dtype=torch.float64, device=X.device)
        grad = torch.empty(b, d, dtype=torch.float64, device=X.device) if need else None
        ones = state.ones(b) if need else None
        ws = state.workspace(b * mc, d, b)
        est = capi.PRB_EST_REINFORCE_MA if grad_estimator == "reinforce_ma" else capi.PRB_EST_REINFORCE
        with torch.cuda.device(X.device):
            st = torch.cuda.current_stream().cuda_stream
            capi.check(lib.prb_mc_value_grad(state.handle, C.byref(sp), C.byref(aq), theta.data_ptr(), b, mc,
                                             capi.ptr(u_int), capi.ptr(u_cat), 0, 0, est, ma_counter.data_ptr(), 1,
                                             capi.ptr(ones), value.data_ptr(), capi.ptr(grad), None, ws.data_ptr(),
                                             ws.numel(), int(state.chunk_cols), st), "prb_mc_value_grad")
        ctx.unit_grad = grad
        return value

    @staticmethod
    def backward(ctx, grad_output):
        none11 = (None,) * 11
        if ctx.unit_grad is None:
            return (None,) + none11
        # the library evaluated the estimator for grad_output = 1; it is linear in grad_output (:613-633)
        g = (ctx.unit_grad * grad_output.unsqueeze(-1)).unsqueeze(-2)
        return (g,) + none11
