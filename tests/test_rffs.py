"""Host side of the Thompson-sampling path: the random-Fourier-feature construction restated from BoTorch
([3P], absent) is checked against what it must approximate -- E[phi(x) . phi(y)] = k(x, y) -- and the weight posterior
against the ridge-regression normal equations.  CPU only."""
import math

import pytest
import torch

from bo_pr_b200.kernels import MaternKernel, ScaleKernel
from bo_pr_b200.rffs import RandomFourierFeatures, get_weights_posterior


def _matern(nu, r):
    if nu == 0.5:
        return torch.exp(-r)
    if nu == 1.5:
        return (1 + math.sqrt(3) * r) * torch.exp(-math.sqrt(3) * r)
    return (1 + math.sqrt(5) * r + 5.0 / 3.0 * r ** 2) * torch.exp(-math.sqrt(5) * r)


@pytest.mark.parametrize("nu", [0.5, 1.5, 2.5])
def test_feature_inner_products_approximate_the_matern_kernel(nu):
    torch.manual_seed(0)
    d, m = 4, 60000
    if nu == 2.5:
        base = MaternKernel(nu=nu, ard_num_dims=d)
    else:  # duck-typed GPyTorch kernel (class name, nu, lengthscale, active_dims): the package's own class is nu = 2.5 only
        base = type("MaternKernel", (), {"nu": nu, "active_dims": None})()
    base.lengthscale = torch.tensor([[0.4, 0.7, 1.1, 0.9]], dtype=torch.float64)
    kernel = ScaleKernel(base)
    kernel.outputscale = torch.tensor(1.7, dtype=torch.float64)
    basis = RandomFourierFeatures(kernel, input_dim=d, num_rff_features=m)
    X = torch.rand(12, d, dtype=torch.float64)
    phi = basis(X)
    approx = phi @ phi.T
    r = ((X.unsqueeze(1) - X.unsqueeze(0)) / base.lengthscale.reshape(-1)).norm(dim=-1)
    exact = 1.7 * _matern(nu, r)
    assert (approx - exact).abs().max() < 0.05, (approx - exact).abs().max()


def test_active_dims_get_zero_rows_and_non_matern_kernels_raise():
    torch.manual_seed(1)
    base = MaternKernel(nu=2.5, ard_num_dims=2, active_dims=[0, 3])
    base.lengthscale = torch.tensor([[0.5, 2.0]], dtype=torch.float64)
    basis = RandomFourierFeatures(base, input_dim=5, num_rff_features=16)
    assert basis.weights.shape == (5, 16)
    assert bool((basis.weights[[1, 2, 4]] == 0).all()) and bool((basis.weights[[0, 3]] != 0).all())
    assert basis.scale == math.sqrt(2.0 / 16)
    prod = MaternKernel(nu=2.5, ard_num_dims=2, active_dims=[0, 1]) * MaternKernel(nu=2.5, active_dims=[2])
    with pytest.raises(NotImplementedError):  # the reference's mixed product kernels fail in BoTorch's RFF as well
        RandomFourierFeatures(ScaleKernel(prod), input_dim=3, num_rff_features=8)


def test_weight_posterior_is_the_ridge_solution():
    torch.manual_seed(2)
    Phi = torch.randn(30, 9, dtype=torch.float64)
    y = torch.randn(30, dtype=torch.float64)
    s2 = 0.37
    mean, L = get_weights_posterior(Phi, y, s2)
    A = Phi.T @ Phi + s2 * torch.eye(9, dtype=torch.float64)
    assert torch.allclose(A @ mean, Phi.T @ y, rtol=1e-10, atol=1e-12)
    assert torch.allclose(L @ L.T, s2 * torch.linalg.inv(A), rtol=1e-9, atol=1e-12)
