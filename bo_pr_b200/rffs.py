"""Random-Fourier-feature posterior samples for Thompson sampling (label ``pr__ts``).

The reference draws ONE approximate GP posterior sample ``f(x) = phi(x)^T w`` and maximises ``PosteriorMean`` of that
deterministic model under PR (``discrete_mixed_bo/rffs.py:24-133``, ``experiment_utils.py:481-492``).  Drawing the sample
(features, Bayesian linear regression on ``phi(X_train)``: one ``m x m`` Cholesky per BO iteration) is the step *before*
the hot path and stays in torch; evaluating ``f`` and its gradient for every (restart, MC sample) inside the PR acquisition
is the hot part and runs in ``libprb200`` (``prb_rff_model_create`` + ``rff_eval_kernel``).

``RandomFourierFeatures`` / ``get_weights_posterior`` restate ``botorch.utils.gp_sampling`` ([3P], BoTorch is not in this
image: recalled from the 0.6-0.8 sources, parity unpinned): ``phi(x) = sqrt(2 s / m) cos((x / l) @ omega + b)``,
``omega ~ N(0, I)`` (RBF) or ``N(0, I) / sqrt(Gamma(nu, nu))`` per feature (Matern), ``b ~ U(0, 2 pi)``; like BoTorch, only
``ScaleKernel(Matern | RBF)`` or a bare Matern / RBF kernel is accepted -- the reference's product kernels raise there too.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import Tensor

from . import _capi as capi
from .models import MIN_VARIANCE, Model, RFFState, _Posterior


class RandomFourierFeatures(torch.nn.Module):
    """``botorch.utils.gp_sampling.RandomFourierFeatures`` for an unbatched kernel.  Random draws use torch's global CPU
    generator in BoTorch's order (``randn`` weights, Gamma mixing for Matern, ``rand`` bias) and are then moved to
    ``device``."""

    def __init__(self, kernel, input_dim: int, num_rff_features: int, sample_shape: Optional[torch.Size] = None,
                 device: Optional[torch.device] = None):
        super().__init__()
        if sample_shape is not None and len(sample_shape) > 0:
            raise NotImplementedError("batches of sampled functions (n_samples > 1) are outside the B200 hot path")
        if type(kernel).__name__ == "ScaleKernel":
            base = kernel.base_kernel
            outputscale = float(torch.as_tensor(kernel.outputscale).detach().reshape(-1)[0])
        else:
            base, outputscale = kernel, 1.0
        kind = type(base).__name__
        if kind not in ("MaternKernel", "RBFKernel"):
            raise NotImplementedError("Only Matern and RBF kernels are supported.")
        ls = torch.as_tensor(base.lengthscale).detach().reshape(-1).to(torch.float64).cpu()
        active = getattr(base, "active_dims", None)
        dims = list(range(input_dim)) if active is None else [int(i) for i in active]
        if ls.numel() == 1:
            ls = ls.expand(len(dims))
        if ls.numel() != len(dims):
            raise ValueError("lengthscale / active_dims mismatch")
        weights = torch.randn(len(dims), num_rff_features, dtype=torch.float64)
        if kind == "MaternKernel":
            nu = float(getattr(base, "nu", 2.5))
            gamma = torch.distributions.Gamma(torch.tensor(nu, dtype=torch.float64),
                                              torch.tensor(nu, dtype=torch.float64)).sample(torch.Size([1, num_rff_features]))
            weights = torch.rsqrt(gamma) * weights
        bias = 2.0 * math.pi * torch.rand(num_rff_features, dtype=torch.float64)
        # x / l @ omega  ==  x @ (omega / l): the lengthscales are folded into the frequency matrix; dims the kernel does
        # not see get zero rows
        W = torch.zeros(input_dim, num_rff_features, dtype=torch.float64)
        W[torch.as_tensor(dims, dtype=torch.long)] = weights / ls.unsqueeze(-1)
        self.register_buffer("weights", W.to(device))
        self.register_buffer("bias", bias.to(device))
        self.outputscale = outputscale
        self.num_rff_features = int(num_rff_features)

    @property
    def scale(self) -> float:
        return math.sqrt(2.0 * self.outputscale / self.num_rff_features)

    def forward(self, X: Tensor) -> Tensor:
        return self.scale * torch.cos(X.to(self.weights) @ self.weights + self.bias)


def get_weights_posterior(X: Tensor, y: Tensor, sigma_sq: float):
    """Posterior ``N(m, S)`` of the Bayesian linear model ``y = X w + eps`` with a standard-normal prior on ``w``
    (``botorch.utils.gp_sampling.get_weights_posterior``): ``A = X^T X + sigma^2 I``, ``m = A^-1 X^T y``,
    ``S = sigma^2 A^-1``.  Returns ``(mean [m], scale_tril [m, m])``."""
    X = X.to(torch.float64)
    A = X.transpose(-2, -1) @ X + sigma_sq * torch.eye(X.shape[-1], dtype=torch.float64, device=X.device)
    L_A = torch.linalg.cholesky(A)
    A_inv = torch.cholesky_solve(torch.eye(X.shape[-1], dtype=torch.float64, device=X.device), L_A)
    mean = (A_inv @ X.transpose(-2, -1) @ y.to(X).unsqueeze(-1)).squeeze(-1)
    cov = A_inv * sigma_sq
    L = torch.linalg.cholesky(0.5 * (cov + cov.transpose(-2, -1)))
    return mean, L


class RFFSampleModel(Model):
    """The deterministic model ``get_gp_samples`` returns (BoTorch ``GenericDeterministicModel`` surface: ``posterior(X).mean``,
    ``num_outputs``).  ``PosteriorMean(model=...)`` over it is the Thompson-sampling acquisition function; its ``state`` is
    what the PR classes hand to the library."""

    num_outputs = 1

    def __init__(self, basis: RandomFourierFeatures, weights: Tensor):
        super().__init__()
        self.basis = basis
        self.register_buffer("weights", weights.detach().to(torch.float64))
        self._state: Optional[RFFState] = None

    @property
    def state(self) -> RFFState:
        if self._state is None:
            self._state = RFFState(self.basis.weights, self.basis.bias, self.weights, self.basis.scale)
        return self._state

    def forward(self, X: Tensor) -> Tensor:
        return self.posterior(X).mean

    def posterior(self, X: Tensor, **kwargs) -> _Posterior:
        desc = capi.PrbAcqDesc(kind=capi.PRB_ACQ_MEAN, ei_variant=0, param=0.0, min_variance=MIN_VARIANCE,
                               acq_var_floor=0.0)
        _, mean, var, _ = self.state.acq_points(desc, X, want_post=True)
        shape = X.shape[:-1]
        return _Posterior(mean.reshape(*shape, 1), var.reshape(*shape, 1))


def get_gp_samples(model, num_outputs: int = 1, n_samples: int = 1, num_rff_features: int = 512) -> RFFSampleModel:
    """One approximate posterior sample of a single-output exact GP (``rffs.py:24-107`` with ``num_outputs == 1`` and
    ``n_samples == 1``, the only use at ``experiment_utils.py:481-487``).  As in BoTorch, the prior mean is ignored:
    the linear model is fitted to the raw ``train_targets``."""
    if num_outputs != 1 or hasattr(model, "models"):
        raise NotImplementedError("multi-output models are outside the B200 hot path")
    if n_samples != 1:
        raise NotImplementedError("batches of sampled functions (n_samples > 1) are outside the B200 hot path")
    train_X = model.train_inputs[0]
    capi.require_cuda(train_X, "train_X")
    basis = RandomFourierFeatures(kernel=model.covar_module, input_dim=train_X.shape[-1],
                                  num_rff_features=num_rff_features, device=train_X.device)
    phi_X = basis(train_X.to(torch.float64))
    sigma_sq = float(torch.as_tensor(model.likelihood.noise).detach().to(torch.float64).mean())
    mean, L = get_weights_posterior(phi_X, model.train_targets.to(phi_X), sigma_sq)
    eps = torch.randn(mean.shape, dtype=torch.float64).to(mean.device)  # MultivariateNormal(loc, scale_tril).sample()
    return RFFSampleModel(basis, mean + L @ eps)


def get_gp_sample_w_transforms(model, num_outputs: int = 1, n_samples: int = 1,
                               num_rff_features: int = 512) -> RFFSampleModel:
    """``rffs.py:110-133``.  Model-level input / outcome transforms are not supported by the B200 path (raise)."""
    for attr in ("input_transform", "outcome_transform"):
        if getattr(model, attr, None) is not None:
            raise NotImplementedError(f"model.{attr} is not supported by the B200 hot path")
    return get_gp_samples(model, num_outputs=num_outputs, n_samples=n_samples, num_rff_features=num_rff_features)
