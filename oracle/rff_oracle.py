"""ORACLE (test infrastructure only): CPU restatement of the deterministic random-Fourier-feature sample the reference
optimises for Thompson sampling.

``discrete_mixed_bo/rffs.py:24-107`` builds ``f(x) = phi(x)^T w`` with BoTorch's ``RandomFourierFeatures`` ([3P], absent
here: ``phi(x) = sqrt(2 s / m) cos((x / l) @ omega + b)``) and wraps it in a ``GenericDeterministicModel``;
``experiment_utils.py:488-492`` puts ``PosteriorMean`` on top.  This module restates the *evaluation* of such a sample in
plain torch float64 on the CPU, differentiable by autograd, with the ``posterior(X) -> (mean, variance)`` surface
``oracle/gp_oracle.py::OraclePosteriorMean`` expects -- so ``pr_oracle.mc_pr`` / ``analytic_pr`` run over it unchanged.
Parity status: evaluation is exact mathematics (pinned by construction); the random construction of (omega, b, w) lives in
the product's host code and is checked against the kernel it approximates (tests/test_rffs.py), not against BoTorch:
parity unpinned for that part.
"""
from __future__ import annotations

import torch
from torch import Tensor


class OracleRFFSample(torch.nn.Module):
    """``f(X) = scale * cos(X @ W + bias) @ weights``; ``W [d, m]`` already divided by the lengthscales."""

    def __init__(self, W: Tensor, bias: Tensor, weights: Tensor, scale: float):
        super().__init__()
        self.W = W.detach().to(torch.float64).cpu()
        self.bias = bias.detach().to(torch.float64).cpu().reshape(-1)
        self.weights = weights.detach().to(torch.float64).cpu().reshape(-1)
        self.scale = float(scale)

    def posterior(self, X: Tensor):
        """``X [..., 1, d] -> (mean [...], variance [...])`` (q = 1 squeezed, as ``OracleGP.posterior``)."""
        x = X.to(torch.float64).squeeze(-2)
        mean = self.scale * (torch.cos(x @ self.W + self.bias) @ self.weights)
        return mean, torch.zeros_like(mean)
