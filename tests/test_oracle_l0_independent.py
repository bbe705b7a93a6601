"""L0 oracle (oracle/gp_oracle.py: Matern-5/2 kernels, exact-GP posterior, EI / UCB) against INDEPENDENT third-party
implementations available in this image: scikit-learn's GaussianProcessRegressor with Matern(nu=2.5) kernels and
scipy.stats.norm.  BoTorch / GPyTorch (the libraries the reference actually calls) are not installed anywhere, so this is
not a pin on their outputs -- it pins the published formulas the oracle restates (kernel, posterior mean / variance, EI)
on an implementation we did not write.  BoTorch-specific constants (variance floors, min_fixed_noise) remain as documented
parameters.  CPU only."""
import math

import numpy as np
import pytest
import torch
from scipy.stats import norm
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import ConstantKernel, Matern

from oracle import gp_oracle as G
from tests.golden_util import load, oracle_spec

BIG = 1e12  # "inactive" dimension of an anisotropic sklearn Matern: (dx / 1e12)^2 adds < 1e-24 to r^2


def _sk_kernel(spec, d):
    assert len(spec) == 1 and not spec[0].additive
    k = ConstantKernel(spec[0].outputscale, constant_value_bounds="fixed")
    for f in spec[0].factors:
        assert f.kind == "matern52" and f.power == 1
        ls = np.full(d, BIG)
        ls[f.dims] = f.lengthscale
        k = k * Matern(length_scale=ls, length_scale_bounds="fixed", nu=2.5)
    return k


@pytest.mark.parametrize("name", ["ackley13_ei", "rosenbrock10_ucb", "mixed_int_f1_ei", "neg_int_onehot_ei"])
def test_posterior_matches_sklearn_gpr(name):
    g = load(name)
    spec = oracle_spec(g)
    X = g["train_X_model"].numpy()
    y = g["train_Y"].numpy()
    d = X.shape[1]
    gpr = GaussianProcessRegressor(kernel=_sk_kernel(spec, d), alpha=g["noise"], optimizer=None, normalize_y=False)
    gpr.fit(X, y - g["mean_const"])  # sklearn has a zero prior mean: fit the residual of the constant mean
    rng = np.random.default_rng(0)
    Xt = X[rng.integers(0, X.shape[0], 50)].copy()
    cont = g["cont_dims"]
    Xt[:, cont] = np.clip(Xt[:, cont] + 0.2 * rng.standard_normal((50, len(cont))), 0, 1)
    mu_sk, sd_sk = gpr.predict(Xt, return_std=True)
    ogp = G.OracleGP(g["train_X_model"], g["train_Y"], spec, g["noise"], g["mean_const"])
    mu_o, var_o = ogp.posterior(torch.from_numpy(Xt))
    K_o = G.eval_kernel(spec, torch.from_numpy(Xt), g["train_X_model"]).numpy()
    np.testing.assert_allclose(K_o, gpr.kernel_(Xt, X), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(mu_o.numpy(), mu_sk + g["mean_const"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(var_o.numpy(), sd_sk ** 2, rtol=1e-6, atol=1e-8)


def test_expected_improvement_and_ucb_closed_forms():
    g = load("ackley13_ei")
    spec = oracle_spec(g)
    ogp = G.OracleGP(g["train_X_model"], g["train_Y"], spec, g["noise"], g["mean_const"])
    X = torch.rand(64, 1, 13, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    X[..., 3:] = X[..., 3:].round()
    mean, var = ogp.posterior(X)
    best = float(g["train_Y"].max())
    sigma = np.sqrt(np.maximum(var.numpy(), 1e-9))
    u = (mean.numpy() - best) / sigma
    ei_ref = sigma * (norm.pdf(u) + u * norm.cdf(u))
    for variant in ("erf", "erfc"):
        ei = G.OracleEI(ogp, best_f=best, variant=variant)(X).numpy()
        np.testing.assert_allclose(ei, ei_ref, rtol=1e-9, atol=1e-15)
    beta = 0.2 * 13 * math.log(2.0)
    ucb = G.OracleUCB(ogp, beta=beta)(X).numpy()
    np.testing.assert_allclose(ucb, mean.numpy() + np.sqrt(beta * var.numpy()), rtol=1e-13)


def test_posterior_interpolates_training_data():
    """Known-answer check: at a training input the posterior mean is y up to the noise level and the
    variance collapses to the noise scale."""
    g = load("rosenbrock10_ucb")
    ogp = G.OracleGP(g["train_X_model"], g["train_Y"], oracle_spec(g), g["noise"], g["mean_const"])
    mean, var = ogp.posterior(g["train_X_model"])
    assert float((mean - g["train_Y"]).abs().max()) < 1e-4
    assert float(var.max()) < 1e-5


def test_categorical_kernel_known_answers():
    """botorch CategoricalKernel: exp(-mean_j([x_j != x'_j] / l_j)) -- by hand."""
    spec = [G.Term(2.0, [G.Factor("categorical", [1, 2], [0.5, 2.0])])]
    x1 = torch.tensor([[0.3, 1.0, 4.0], [0.9, 1.0, 3.0], [0.1, 2.0, 3.0]], dtype=torch.float64)
    x2 = torch.tensor([[0.0, 1.0, 4.0]], dtype=torch.float64)
    k = G.eval_kernel(spec, x1, x2).reshape(-1)
    expect = [2.0, 2.0 * math.exp(-(0 + 1 / 2.0) / 2), 2.0 * math.exp(-(1 / 0.5 + 1 / 2.0) / 2)]
    np.testing.assert_allclose(k.numpy(), expect, rtol=1e-15)
