"""Synthetic PR-EI workload of BASELINE.json's sweep, shared by bench.py and the tests.

d = 13 laid out as the reference orders columns (continuous first): 3 continuous in [0, 1] + 10 binary (the ackley13
shape); kernel "mixed": outputscale 1.3, ARD lengthscales U[0.3, 0.8] on the continuous dims, isotropic 1.2 on the
binary dims, constant mean 0.1, fixed noise 1e-6; y = a draw from the GP prior at X_train, standardised;
best_f = max y.  theta: restarts ~ U[0, 1]^{b x 1 x 13}.  Everything is seeded.

This module only SYNTHESISES inputs (it is not on the product path): the prior draw needs K(X, X) once, formed here with
a few lines of plain torch on whatever device the caller passes.
"""
from __future__ import annotations

import math

import torch

D = 13
D_CONT = 3
BINARY_DIMS = list(range(3, 13))


def _matern52(x1, x2, ls):
    d = (x1 / ls).unsqueeze(-2) - (x2 / ls).unsqueeze(-3)
    r = d.pow(2).sum(-1).clamp_min(1e-30).sqrt()
    s5 = math.sqrt(5.0)
    return (1.0 + s5 * r + (5.0 / 3.0) * r * r) * torch.exp(-s5 * r)


def make_problem(n: int, seed: int = 0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, D, generator=g, dtype=torch.float64)
    X[:, D_CONT:] = X[:, D_CONT:].round()
    cont_ls = (0.3 + 0.5 * torch.rand(D_CONT, generator=g, dtype=torch.float64)).tolist()
    hp = dict(cont_lengthscale=cont_ls, binary_lengthscale=[1.2], categorical_lengthscale=[], outputscale=1.3,
              outputscale_sum=1.0)
    z = torch.randn(n, generator=g, dtype=torch.float64)
    Xd = X.to(device)
    K = torch.empty(n, n, dtype=torch.float64, device=device)
    ls_c = torch.tensor(cont_ls, dtype=torch.float64, device=device)
    ls_b = torch.full((len(BINARY_DIMS),), 1.2, dtype=torch.float64, device=device)
    for s in range(0, n, 512):
        e = min(n, s + 512)
        K[s:e] = 1.3 * _matern52(Xd[s:e, :D_CONT], Xd[:, :D_CONT], ls_c) * _matern52(Xd[s:e, D_CONT:], Xd[:, D_CONT:], ls_b)
    K.diagonal().add_(1e-6)
    y = (torch.linalg.cholesky(K) @ z.to(device)).cpu()
    if n > 1:
        y = (y - y.mean()) / y.std()
    case = dict(dim=D, integer_indices=BINARY_DIMS, integer_bounds=[[0.0] * 10, [1.0] * 10], categorical_features=None,
                kernel_type="mixed", binary_dims=BINARY_DIMS, acq="ei", apply_numeric=False,
                grad_estimator="reinforce_ma", beta=0.2 * D * math.log(2.0))
    return dict(case=case, train_X_model=X, train_Y=y, hp=hp, noise=1e-6, mean_const=0.1, cont_dims=[0, 1, 2],
                cat_dims=[])


def make_theta(b: int, seed: int = 1):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(b, 1, D, generator=g, dtype=torch.float64)


def sobol_base_samples(mc: int, n_cols: int = 10, seed: int = 1234):
    """[mc, 1, n_cols] scrambled Sobol, as the reference draws them (input.py:667-672)."""
    from torch.quasirandom import SobolEngine
    return SobolEngine(n_cols, scramble=True, seed=seed).draw(mc, dtype=torch.float64).view(mc, 1, n_cols)


GENERIC_CASES = {
    # rosenbrock10_scaled shape: 4 continuous + 6 ordinals {0..3}, ALL inside one ARD Matern -> the
    # kernel does not separate into (per-restart) x (per-sample) factors: generic path
    "rosenbrock": dict(dim=10, integer_indices=list(range(4, 10)), integer_bounds=[[0.0] * 6, [3.0] * 6],
                       categorical_features=None, kernel_type="mixed", binary_dims=[], acq="ucb",
                       beta=0.2 * 10 * math.log(2 * 100), apply_numeric=False, grad_estimator="reinforce_ma"),
    # chemistry shape, MC variant: 2 continuous + categoricals (4, 12, 4), kernel "mixed_categorical" (two additive terms,
    # squared categorical factor), base AF sees the numeric layout (OneHotToNumeric)
    "chemistry": dict(dim=22, integer_indices=[], integer_bounds=[[], []], categorical_features={2: 4, 3: 12, 4: 4},
                      kernel_type="mixed_categorical", binary_dims=[], acq="ei", apply_numeric=True,
                      grad_estimator="reinforce_ma"),
}


def make_generic_problem(kind: str, n: int, restarts: int, mc: int, seed: int = 0):
    """(problem, theta [restarts, 1, dim]) for the kernel structures that do NOT separate, at any n_train.  X_train is laid
    out as the model sees it (numeric layout when apply_numeric); y is a smooth seeded function of it, standardised (any
    targets give a valid exact-GP state; the hot path only sees alpha and the Cholesky factor)."""
    c = dict(GENERIC_CASES[kind], mc=mc, n_train=n, b=restarts, seed=seed)
    g = torch.Generator().manual_seed(seed + 77)
    cf = c["categorical_features"] or {}
    ints = c["integer_indices"]
    ib = torch.tensor(c["integer_bounds"], dtype=torch.float64).reshape(2, -1)
    n_cont = min(cf.keys()) if cf else c["dim"] - len(ints)
    if cf:
        n_cont -= len(ints)
    cols = [torch.rand(n, n_cont, generator=g, dtype=torch.float64)]
    for k in range(len(ints)):
        lo, hi = float(ib[0, k]), float(ib[1, k])
        v = torch.randint(int(lo), int(hi) + 1, (n, 1), generator=g).to(torch.float64)
        cols.append((v - lo) / (hi - lo))
    for card in cf.values():  # numeric codes (apply_numeric) -- the only layout the categorical cases here use
        cols.append(torch.randint(0, card, (n, 1), generator=g).to(torch.float64))
    Xk = torch.cat(cols, dim=-1)
    dk = Xk.shape[-1]
    cat_dims = list(cf.keys()) if c["kernel_type"] == "mixed_categorical" else []
    cont_dims = sorted(set(range(dk)) - set(c["binary_dims"]) - set(cat_dims))
    hp = dict(cont_lengthscale=(0.4 + 0.8 * torch.rand(len(cont_dims), generator=g, dtype=torch.float64)).tolist(),
              binary_lengthscale=[], categorical_lengthscale=(0.6 + 1.2 * torch.rand(len(cat_dims), generator=g,
                                                                                      dtype=torch.float64)).tolist(),
              outputscale=1.1, outputscale_sum=0.4)
    w = torch.randn(dk, generator=g, dtype=torch.float64)
    y = torch.sin(3.0 * (Xk / Xk.abs().max(dim=0).values.clamp_min(1.0)) @ w) + 0.1 * torch.randn(n, generator=g,
                                                                                                    dtype=torch.float64)
    y = (y - y.mean()) / y.std()
    theta = torch.rand(restarts, 1, c["dim"], generator=g, dtype=torch.float64)
    return dict(case=c, train_X_model=Xk, train_Y=y, hp=hp, noise=1e-4, mean_const=0.05, cont_dims=cont_dims,
                cat_dims=cat_dims, base_samples=None, base_samples_categorical=None), theta
