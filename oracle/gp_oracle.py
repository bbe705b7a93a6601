"""L0 oracle: exact-GP posterior, mixed kernel and base acquisition functions (torch CPU, float64).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **Parity unpinned**: this arithmetic lives in
BoTorch / GPyTorch (third party, unpinned in the reference's ``setup.py:9-20``, absent from this image);
the formulas below are the libraries' published ones as recorded in SURVEY.md App. B, anchored on the
reference's call sites:

* kernel structure       ``discrete_mixed_bo/kernels.py:18-101`` (``get_kernel``)
* model                  ``discrete_mixed_bo/experiment_utils.py:237-351`` (``FixedNoiseGP``, constant mean)
* EI / UCB / mean        ``discrete_mixed_bo/experiment_utils.py:391-393, 476-480, 489-492``
* where L1 calls into it ``discrete_mixed_bo/probabilistic_reparameterization.py:314, 491, 540``

Everything is written so torch autograd can differentiate it, because the reference obtains the pathwise
gradient for the continuous parameters with ``torch.autograd.grad`` through exactly this computation
(``probabilistic_reparameterization.py:624-633``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor

SQRT5 = math.sqrt(5.0)


# ----------------------------------------------------------------------------------------------------------------------
# Kernel specification (plain data; the product has its own classes, tests translate between the two)
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class Factor:
    """One multiplicative factor.  kind: "matern52" (gpytorch MaternKernel nu=2.5) or "categorical"
    (botorch CategoricalKernel).  ``lengthscale`` has one entry per active dim (isotropic = repeated)."""

    kind: str
    dims: List[int]
    lengthscale: List[float]
    power: int = 1


@dataclass
class Term:
    """outputscale * prod_f factor_f ** power_f  (gpytorch ScaleKernel(ProductKernel(...)))."""

    outputscale: float
    factors: List[Factor] = field(default_factory=list)
    # additive inner structure: if True the factors are *summed* (ScaleKernel(AdditiveKernel(...)))
    additive: bool = False


KernelSpec = List[Term]


def reference_kernel_spec(
    kernel_type: str,
    dim: int,
    binary_dims: Sequence[int],
    categorical_transformed_features: Dict[int, int],
    cont_lengthscale: Sequence[float],
    binary_lengthscale: Sequence[float] = (),
    categorical_lengthscale: Sequence[float] = (),
    outputscale: float = 1.0,
    outputscale_sum: float = 1.0,
    use_ard_binary: bool = False,
    latent_lengthscale: Sequence[float] = (),
) -> KernelSpec:
    """Structure of ``get_kernel`` (``kernels.py:18-101``) with explicit hyperparameters.

    * ``kernels.py:29-37``  categorical dims: the numeric columns (mixed_categorical) or the one-hot tail
    * ``kernels.py:39-42``  continuous dims = everything that is not binary (and not categorical if
      ``mixed_categorical``) -- non-binary ordinals and, for kernel_type "mixed", one-hot columns too
    * ``kernels.py:45-63``  Matern-5/2 ARD on continuous dims, Matern-5/2 (isotropic unless
      ``use_ard_binary``) on binary dims
    * ``kernels.py:64-72``  CategoricalKernel (ARD) for ``mixed_categorical``
    * ``kernels.py:85-89``  "mixed": Scale(k0 * k1 * ...)
    * ``kernels.py:90-94``  "mixed_categorical": the second loop multiplies ``prod_kernel`` by ``kernels[1:]``
      AGAIN, so every factor but the first is squared:  Scale(k0 * k1^2 * ...) + Scale(k0 + k1 + ...)
      (SURVEY.md App. C.1 -- replicated, not fixed).
    * ``kernels.py:73-84``  "mixed_latent": ``categorical_transformed_features = {first latent column: latent dim}`` over
      the LatentCategoricalEmbedding layout; categorical dims = everything from the first latent column on; one isotropic
      Matern-5/2 per set of latent columns (``latent_lengthscale``: one value per categorical); Scale(product).
    """
    if kernel_type not in ("mixed", "mixed_categorical", "mixed_latent"):
        raise ValueError(f"{kernel_type} is not supported by the oracle.")
    if kernel_type == "mixed_categorical":
        categorical_dims = list(categorical_transformed_features.keys())
    elif len(categorical_transformed_features) > 0:
        categorical_dims = list(range(min(categorical_transformed_features.keys()), dim))
    else:
        categorical_dims = []
    cont_dims = sorted(set(range(dim)) - set(binary_dims))
    if kernel_type in ("mixed_categorical", "mixed_latent"):
        cont_dims = sorted(set(cont_dims) - set(categorical_dims))
    factors: List[Factor] = []
    if len(cont_dims) > 0:
        assert len(cont_lengthscale) == len(cont_dims)
        factors.append(Factor("matern52", list(cont_dims), [float(v) for v in cont_lengthscale]))
    if len(binary_dims) > 0:
        if use_ard_binary:
            assert len(binary_lengthscale) == len(binary_dims)
            ls = [float(v) for v in binary_lengthscale]
        else:
            assert len(binary_lengthscale) == 1
            ls = [float(binary_lengthscale[0])] * len(binary_dims)
        factors.append(Factor("matern52", list(binary_dims), ls))
    if kernel_type == "mixed_categorical" and len(categorical_dims) > 0:
        assert len(categorical_lengthscale) == len(categorical_dims)
        factors.append(Factor("categorical", list(categorical_dims), [float(v) for v in categorical_lengthscale]))
    if kernel_type == "mixed_latent":
        assert len(latent_lengthscale) == len(categorical_transformed_features)
        for (start, latent_dim), ls in zip(categorical_transformed_features.items(), latent_lengthscale):
            factors.append(Factor("matern52", list(range(start, start + latent_dim)), [float(ls)] * latent_dim))
    if kernel_type in ("mixed", "mixed_latent"):
        return [Term(float(outputscale), [Factor(f.kind, f.dims, f.lengthscale, 1) for f in factors])]
    prod = [Factor(f.kind, f.dims, f.lengthscale, 1 if i == 0 else 2) for i, f in enumerate(factors)]
    summ = [Factor(f.kind, f.dims, f.lengthscale, 1) for f in factors]
    return [Term(float(outputscale), prod), Term(float(outputscale_sum), summ, additive=True)]


# ----------------------------------------------------------------------------------------------------------------------
# Kernel evaluation  [3P] gpytorch MaternKernel / ScaleKernel / ProductKernel / AdditiveKernel, botorch CategoricalKernel
# ----------------------------------------------------------------------------------------------------------------------
def _matern52(x1: Tensor, x2: Tensor, ls: Tensor) -> Tensor:
    """gpytorch MaternKernel(nu=2.5): inputs divided by the lengthscale, Euclidean distance clamped at
    sqrt(1e-30), k = (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r)  (SURVEY.md App. B.1).  GPyTorch subtracts the
    column mean of x1 from both inputs first (a translation) and forms r^2 by an expanded GEMM; the direct
    difference used here agrees to ~1e-15."""
    a = (x1 / ls).unsqueeze(-2)  # [..., B, 1, k]
    b = (x2 / ls).unsqueeze(-3)  # [..., 1, n, k]
    sq = ((a - b) ** 2).sum(-1)
    r = sq.clamp_min(1e-30).sqrt()
    return (1.0 + SQRT5 * r + (5.0 / 3.0) * r * r) * torch.exp(-SQRT5 * r)


def _categorical(x1: Tensor, x2: Tensor, ls: Tensor) -> Tensor:
    """botorch CategoricalKernel: exp(-mean_j(1[x_j != x'_j] / l_j))  (SURVEY.md App. B.2)."""
    delta = (x1.unsqueeze(-2) != x2.unsqueeze(-3)).to(x1.dtype)
    return torch.exp(-(delta / ls).mean(-1))


def eval_kernel(spec: KernelSpec, x1: Tensor, x2: Tensor) -> Tensor:
    """K(x1[..., B, d], x2[n, d]) -> [..., B, n]."""
    out = None
    for term in spec:
        acc = None
        for f in term.factors:
            dims = torch.as_tensor(f.dims, dtype=torch.long)
            ls = torch.as_tensor(f.lengthscale, dtype=x1.dtype)
            a, b = x1[..., dims], x2[..., dims]
            k = _matern52(a, b, ls) if f.kind == "matern52" else _categorical(a, b, ls)
            k1 = k
            for _ in range(f.power - 1):
                k = k * k1
            if acc is None:
                acc = k
            elif term.additive:
                acc = acc + k
            else:
                acc = acc * k
        acc = term.outputscale * acc
        out = acc if out is None else out + acc
    return out


def kernel_diag(spec: KernelSpec, dtype=torch.float64) -> float:
    """k(x, x): every factor is 1 at zero distance."""
    v = 0.0
    for term in spec:
        v += term.outputscale * (len(term.factors) if term.additive else 1.0)
    return v


# ----------------------------------------------------------------------------------------------------------------------
# Exact GP  [3P] botorch FixedNoiseGP + gpytorch DefaultPredictionStrategy (Cholesky path), SURVEY.md App. B.3
# ----------------------------------------------------------------------------------------------------------------------
def latent_embedding_transform(tables: Sequence[Tensor], categ_idcs: Sequence[int], dim: int):
    """``EmbeddingTransform.transform`` (``input.py:784-802``) with fixed tables: categorical column ``categ_idcs[j]`` (an
    integer code) -> row of ``tables[j]``; output ``[non-categorical columns, embeddings in order]``.  The tables are taken
    as given (``get_emb_table``'s pinning of row 0 / entry [1, 0], ``input.py:902-909``, is applied by the caller)."""
    mask = torch.ones(dim, dtype=torch.bool)
    mask[list(categ_idcs)] = False

    def transform(X: Tensor) -> Tensor:
        parts = [X[..., mask]]
        for idx, table in zip(categ_idcs, tables):
            codes = X[..., idx].reshape(-1).long()
            parts.append(table.to(X).index_select(0, codes).reshape(*X.shape[:-1], table.shape[-1]))
        return torch.cat(parts, dim=-1)

    return transform


class OracleGP(torch.nn.Module):
    """FixedNoiseGP with a constant mean, in eval mode, exact Cholesky caches.

    mean_cache  = (K + diag(noise))^-1 (y - c)          (gpytorch exact_prediction_strategies mean_cache)
    covar_cache = L^-T, L = chol(K + diag(noise))        (explicit inverse root on the Cholesky path)
    mu*  = c + k* . mean_cache ;  var* = k** - || k*^T L^-T ||^2, floored at gpytorch ``min_variance`` (1e-10, f64)
    Posterior is noise-free (``observation_noise=False``), q = 1.
    """

    def __init__(self, train_X: Tensor, train_Y: Tensor, spec: KernelSpec, noise, mean_const: float,
                 min_variance: float = 1e-10, input_transform=None):
        """``input_transform``: callable applied to the inputs of ``posterior`` (a model-level LatentCategoricalEmbedding,
        input.py:777-802); ``train_X`` is given already transformed."""
        super().__init__()
        self.input_transform = input_transform
        train_X = train_X.detach().to(torch.float64).cpu()
        y = train_Y.detach().to(torch.float64).cpu().reshape(-1)
        n = train_X.shape[0]
        noise = torch.as_tensor(noise, dtype=torch.float64).reshape(-1)
        if noise.numel() == 1:
            noise = noise.expand(n)
        self.spec = spec
        self.train_X = train_X
        self.train_Y = y
        self.mean_const = float(mean_const)
        self.noise = noise.clone()
        self.min_variance = float(min_variance)
        K = eval_kernel(spec, train_X, train_X) + torch.diag(self.noise)
        self.L = torch.linalg.cholesky(K)
        self.Linv = torch.linalg.solve_triangular(self.L, torch.eye(n, dtype=torch.float64), upper=False)
        self.alpha = torch.cholesky_solve((y - self.mean_const).unsqueeze(-1), self.L).squeeze(-1)
        self.kss = kernel_diag(spec)

    def posterior(self, X: Tensor):
        """X [..., 1, d] (q = 1) or [..., d] -> (mean [...], variance [...])."""
        if X.shape[-2:-1] == (1,) and X.ndim >= 2:
            Xp = X.squeeze(-2)
        else:
            Xp = X
        if self.input_transform is not None:
            Xp = self.input_transform(Xp)
        ks = eval_kernel(self.spec, Xp, self.train_X)  # [..., n]
        mean = self.mean_const + ks @ self.alpha
        root = ks @ self.Linv.transpose(-1, -2)  # k*^T L^-T
        var = self.kss - (root * root).sum(-1)
        var = var.clamp_min(self.min_variance)
        return mean, var


# ----------------------------------------------------------------------------------------------------------------------
# Base acquisition functions  [3P] botorch.acquisition.analytic, SURVEY.md App. B.4
# ----------------------------------------------------------------------------------------------------------------------
_LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


class OracleAcq(torch.nn.Module):
    """``acq_function(X[b, mc, 1, d]) -> [b, mc]``; duck-types the BoTorch analytic AF surface the reference
    touches (``.model``, ``.best_f`` / ``.beta``, ``.X_baseline``)."""

    def __init__(self, model: OracleGP):
        super().__init__()
        self.model = model
        self.X_pending = None

    @property
    def X_baseline(self):
        return None

    def set_X_pending(self, X_pending=None):
        self.X_pending = X_pending


class OracleEI(OracleAcq):
    """ExpectedImprovement as in BoTorch 0.6-0.8 (the reference's era: FixedNoiseGP / fit_gpytorch_model):
    sigma = sqrt(clamp_min(var, 1e-9)); u = (mean - best_f)/sigma;
    EI = sigma * (exp(Normal.log_prob(u)) + u * Normal.cdf(u)),  cdf = 0.5 (1 + erf(u / sqrt2)).
    ``variant="erfc"`` gives the later (>= 0.9) helper: phi(u) + u * 0.5 erfc(-u / sqrt2)."""

    def __init__(self, model, best_f, var_floor: float = 1e-9, variant: str = "erf"):
        super().__init__(model)
        self.best_f = torch.as_tensor(best_f, dtype=torch.float64)
        self.var_floor = var_floor
        self.variant = variant

    def forward(self, X):
        mean, var = self.model.posterior(X)
        sigma = var.clamp_min(self.var_floor).sqrt()
        u = (mean - self.best_f) / sigma
        updf = torch.exp(-(u ** 2) / 2 - _LOG_SQRT_2PI)
        if self.variant == "erf":
            ucdf = 0.5 * (1 + torch.erf(u / math.sqrt(2)))
        else:
            ucdf = 0.5 * torch.erfc(-u / math.sqrt(2))
        return sigma * (updf + u * ucdf)


class OracleUCB(OracleAcq):
    """UpperConfidenceBound: mean + sqrt(beta * var)   (experiment_utils.py:476-480; beta = 0.2 d ln(2 iter))."""

    def __init__(self, model, beta: float):
        super().__init__(model)
        self.beta = torch.as_tensor(beta, dtype=torch.float64)

    def forward(self, X):
        mean, var = self.model.posterior(X)
        return mean + (self.beta * var).sqrt()


class OraclePosteriorMean(OracleAcq):
    """PosteriorMean (experiment_utils.py:489-492; the reference applies it to an RFF sample for TS)."""

    def forward(self, X):
        mean, _ = self.model.posterior(X)
        return mean


# ----------------------------------------------------------------------------------------------------------------------
# Model lists and constrained EI  [3P] botorch ModelListGP + ConstrainedExpectedImprovement (BoTorch 0.6-0.8),
# call site experiment_utils.py:354-392 (get_EI with num_constraints > 0; welded_beam, pressure_vessel, ...)
# ----------------------------------------------------------------------------------------------------------------------
class OracleModelList(torch.nn.Module):
    """Independent single-output GPs over the same inputs (``ModelListGP(*models)``, experiment_utils.py:337-346):
    ``posterior(X)`` -> (mean [..., m], variance [..., m])."""

    def __init__(self, *models: OracleGP):
        super().__init__()
        self.models = torch.nn.ModuleList(models)

    def posterior(self, X: Tensor):
        ms, vs = zip(*[m.posterior(X) for m in self.models])
        return torch.stack(ms, dim=-1), torch.stack(vs, dim=-1)


class OracleConstrainedEI(OracleAcq):
    """``ConstrainedExpectedImprovement(model, best_f, objective_index, constraints)`` as BoTorch <= 0.8 computes it:
    sigma = sqrt(var).clamp_min(1e-9) for EVERY output (note: the clamp is on sigma here, on the variance in plain EI);
    EI of the objective output times prod_i P(lower_i <= Y_i <= upper_i) with Normal.cdf = 0.5 (1 + erf(x / sqrt2)):
    lower only: 1 - cdf((lower - mean)/sigma); upper only: cdf((upper - mean)/sigma); both: the difference.
    ``constraints = {output index: (lower or None, upper or None)}`` (get_EI passes ``(slack, None)``)."""

    def __init__(self, model: OracleModelList, best_f, objective_index: int, constraints, sigma_floor: float = 1e-9):
        super().__init__(model)
        self.best_f = torch.as_tensor(best_f, dtype=torch.float64)
        self.objective_index = int(objective_index)
        self.constraints = dict(constraints)
        self.sigma_floor = sigma_floor

    def forward(self, X):
        means, variances = self.model.posterior(X)
        sigmas = variances.sqrt().clamp_min(self.sigma_floor)
        mean_obj, sigma_obj = means[..., self.objective_index], sigmas[..., self.objective_index]
        u = (mean_obj - self.best_f) / sigma_obj
        ei_pdf = torch.exp(-(u ** 2) / 2 - _LOG_SQRT_2PI)
        ei_cdf = 0.5 * (1 + torch.erf(u / math.sqrt(2)))
        ei = sigma_obj * (ei_pdf + u * ei_cdf)
        cdf = lambda x: 0.5 * (1 + torch.erf(x / math.sqrt(2)))  # noqa: E731
        prob = torch.ones_like(ei)
        for i, (lo, hi) in self.constraints.items():
            m, s = means[..., i], sigmas[..., i]
            if lo is not None and hi is not None:
                prob = prob * (cdf((hi - m) / s) - cdf((lo - m) / s))
            elif lo is not None:
                prob = prob * (1 - cdf((lo - m) / s))
            elif hi is not None:
                prob = prob * cdf((hi - m) / s)
        return ei * prob


# ----------------------------------------------------------------------------------------------------------------------
# Analytic expected hypervolume improvement  [3P] botorch.acquisition.multi_objective.analytic.ExpectedHypervolumeImprovement
# (BoTorch 0.6-0.8), call site experiment_utils.py:395-402 (get_EHVI; label pr__ehvi).  PARITY UNPINNED: BoTorch is absent
# from this image and the reference repository holds no recorded EHVI values; the formulas below restate the published
# method (Yang et al. 2019 as BoTorch implements it) and are cross-checked against a Monte Carlo estimate of the
# hypervolume improvement in tests/test_ehvi_host.py.
# ----------------------------------------------------------------------------------------------------------------------
def grid_nondominated_cells(pareto_Y: Tensor, ref_point: Tensor) -> Tensor:
    """A box decomposition of the non-dominated region above ``ref_point`` INDEPENDENT of the product's: the full grid
    spanned by every objective's distinct Pareto values (plus ref and +inf), keeping the boxes whose lower corner no Pareto
    point dominates.  ``O((K + 1)^m)`` boxes -- test infrastructure only.  Returns ``[2, n_cells, m]``."""
    import itertools
    m = ref_point.numel()
    P = pareto_Y.reshape(-1, m).to(torch.float64)
    ref = ref_point.to(torch.float64)
    edges = []
    for k in range(m):
        vals = sorted({float(ref[k])} | {float(v) for v in P[:, k] if float(v) > float(ref[k])})
        edges.append(vals + [math.inf])
    lows, highs = [], []
    for idx in itertools.product(*[range(len(e) - 1) for e in edges]):
        lo = torch.tensor([edges[k][i] for k, i in enumerate(idx)], dtype=torch.float64)
        hi = torch.tensor([edges[k][i + 1] for k, i in enumerate(idx)], dtype=torch.float64)
        # the open box (lo, hi) is dominated iff some Pareto point is >= hi everywhere (grid boxes are never split)
        if P.numel() > 0 and bool((P >= hi).all(-1).any()):
            continue
        lows.append(lo)
        highs.append(hi)
    return torch.stack([torch.stack(lows), torch.stack(highs)])


class OracleEHVI(OracleAcq):
    """q = 1 analytic EHVI over a model list, every output an objective (maximised):

        mu, sigma = posterior mean, sqrt(variance.clamp_min(1e-9));  upper cell bounds clamp_max(1e10)
        psi(l, u) = sigma * exp(Normal.log_prob((u - mu)/sigma)) + (mu - l) * (1 - cdf((u - mu)/sigma))
        nu(l, u)  = (u - l) * (1 - cdf((u - mu)/sigma))
        EHVI = sum over cells, sum over the 2^m choices of {psi(l, l) - psi(l, u), nu(l, u)} per objective, of the product.
    """

    def __init__(self, model: OracleModelList, ref_point, cell_bounds: Tensor, var_floor: float = 1e-9):
        super().__init__(model)
        self.ref_point = torch.as_tensor(ref_point, dtype=torch.float64)
        self.cell_lower_bounds = cell_bounds[0].to(torch.float64)
        self.cell_upper_bounds = cell_bounds[1].to(torch.float64)
        self.var_floor = var_floor
        import itertools
        m = self.ref_point.numel()
        self._choices = torch.tensor(list(itertools.product(*[[0, 1] for _ in range(m)])), dtype=torch.long)

    @staticmethod
    def _cdf(x):
        return 0.5 * (1 + torch.erf(x / math.sqrt(2)))

    def _psi(self, lower, upper, mu, sigma):
        u = (upper - mu) / sigma
        return sigma * torch.exp(-(u ** 2) / 2 - _LOG_SQRT_2PI) + (mu - lower) * (1 - self._cdf(u))

    def _nu(self, lower, upper, mu, sigma):
        return (upper - lower) * (1 - self._cdf((upper - mu) / sigma))

    def forward(self, X):
        means, variances = self.model.posterior(X)  # [..., m]
        mu = means.unsqueeze(-2)  # [..., 1, m] against cells [n_cells, m]
        sigma = variances.clamp_min(self.var_floor).sqrt().unsqueeze(-2)
        upper = self.cell_upper_bounds.clamp_max(1e10)
        lower = self.cell_lower_bounds
        psi_lu = self._psi(lower, upper, mu, sigma)
        psi_ll = self._psi(lower, lower, mu, sigma)
        nu = self._nu(lower, upper, mu, sigma)
        stacked = torch.stack([psi_ll - psi_lu, nu], dim=-2)  # [..., n_cells, 2, m]
        m = lower.shape[-1]
        picked = torch.stack([stacked[..., self._choices[:, k], k] for k in range(m)], dim=-1)  # [..., n_cells, 2^m, m]
        return picked.prod(dim=-1).sum(dim=-1).sum(dim=-1)
