"""L1 oracle: restatement of the reference's probabilistic-reparameterization path (torch CPU, float64).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **Pinned**: every function here is checked against
outputs of the reference's own code (imported unmodified from ``/root/reference`` by ``oracle/gen_golden.py``
and stored under ``tests/golden/``) by ``tests/test_oracle_golden.py``.

Paths cited are relative to ``/root/reference/discrete_mixed_bo``.

Column-order contract (``problems/base.py:64-78``, ``probabilistic_reparameterization.py:153-163``):
``X = [continuous..., integers..., one-hot block of categorical 1, one-hot block 2, ...]``;
``categorical_features = {first numeric index of categorical j: cardinality_j}``; one-hot blocks start at the
smallest key and run contiguously to ``dim`` (``input.py:583-594``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import Tensor


@dataclass
class PRSpace:
    """Static description of the mixed search space as the reference's constructors receive it
    (``probabilistic_reparameterization.py:102-112``)."""

    dim: int
    integer_indices: List[int]
    integer_bounds: Tensor  # [2, n_int] (lo, hi) in the un-normalised integer space
    categorical_features: Optional[Dict[int, int]] = None
    tau: float = 0.1

    def __post_init__(self):
        self.integer_indices = list(self.integer_indices or [])
        self.integer_bounds = torch.as_tensor(self.integer_bounds, dtype=torch.float64).reshape(2, -1)
        cf = self.categorical_features or {}
        self.cat_starts: List[int] = []
        self.cat_cards: List[int] = []
        start = None
        for i, card in cf.items():  # input.py:583-594
            if start is None:
                start = i
            self.cat_starts.append(start)
            self.cat_cards.append(card)
            start = start + card
        self.discrete_indices = list(self.integer_indices)
        for s, c in zip(self.cat_starts, self.cat_cards):
            self.discrete_indices += list(range(s, s + c))
        self.categorical_indices = list(range(min(cf.keys()), self.dim)) if len(cf) > 0 else []
        disc = set(self.integer_indices) | set(self.categorical_indices)
        self.cont_indices = sorted(set(range(self.dim)) - disc)  # probabilistic_reparameterization.py:174-181
        self.n_int = len(self.integer_indices)
        self.n_cat = len(self.cat_starts)
        self.n_disc = len(self.discrete_indices)
        if self.n_cat > 0:
            self.numeric_dim = min(self.cat_starts) + self.n_cat  # input.py:65
        else:
            self.numeric_dim = self.dim


# ----------------------------------------------------------------------------------------------------------------------
# un/normalise  (probabilistic_reparameterization.py:47-63, 83-93 -> botorch Normalize on the integer columns only)
# ----------------------------------------------------------------------------------------------------------------------
def unnormalize(space: PRSpace, X: Tensor) -> Tensor:
    """x_un = min + x * range on the integer columns (separate multiply and add)."""
    if space.n_int == 0:
        return X
    X_new = X.clone()
    lo, hi = space.integer_bounds[0], space.integer_bounds[1]
    idx = torch.as_tensor(space.integer_indices)
    X_new[..., idx] = lo + X[..., idx] * (hi - lo)
    return X_new


def normalize(space: PRSpace, X: Tensor) -> Tensor:
    """x = (x_un - min) / range on the integer columns."""
    if space.n_int == 0:
        return X
    X_new = X.clone()
    lo, hi = space.integer_bounds[0], space.integer_bounds[1]
    idx = torch.as_tensor(space.integer_indices)
    X_new[..., idx] = (X[..., idx] - lo) / (hi - lo)
    return X_new


# ----------------------------------------------------------------------------------------------------------------------
# rounding probabilities  (input.py:613-635; identical twin at 385-408)
# ----------------------------------------------------------------------------------------------------------------------
def rounding_prob(space: PRSpace, X_un: Tensor) -> Tensor:
    """ints: sigmoid((|x| - floor|x| - 0.5)/tau) on UN-normalised x; cats: softmax((x - 0.5)/tau) per one-hot
    block.  Returns the ``discrete_indices`` columns: integers first, then all one-hot columns."""
    Xp = X_un.detach().clone()
    if space.n_int > 0:
        idx = torch.as_tensor(space.integer_indices)
        a = Xp[..., idx].abs()
        Xp[..., idx] = torch.sigmoid((a - a.floor() - 0.5) / space.tau)
    for s, c in zip(space.cat_starts, space.cat_cards):
        Xp[..., s:s + c] = torch.softmax((Xp[..., s:s + c] - 0.5) / space.tau, dim=-1)
    return Xp[..., torch.as_tensor(space.discrete_indices, dtype=torch.long)]


# ----------------------------------------------------------------------------------------------------------------------
# MC sampling  (input.py:637-746)
# ----------------------------------------------------------------------------------------------------------------------
def mc_round(space: PRSpace, X_un: Tensor, base_int: Optional[Tensor], base_cat: Optional[Tensor]) -> Tensor:
    """X_un [b, 1, q=1, d] (integer columns un-normalised) -> [b, mc, 1, d].

    * 648      expand over mc, clone
    * 679-688  z = 1[p >= u];  x' = (-1)^{[x<0]} (floor|x| + z)
    * 690-692  clamp to integer_bounds
    * 723-744  categorical: first index with cumsum(p) > u, **0 when none** (SURVEY.md App. C.4); one-hot
    Base uniforms ``base_int [mc, 1, n_int]`` / ``base_cat [mc, 1, n_cat]`` are shared by all b.
    """
    mc = base_int.shape[0] if base_int is not None else base_cat.shape[0]
    X_exp = X_un.expand(*X_un.shape[:-3], mc, *X_un.shape[-2:]).clone()
    P = rounding_prob(space, X_un)
    if space.n_int > 0:
        idx = torch.as_tensor(space.integer_indices)
        X_int = X_un[..., idx].detach()
        a = X_int.abs()
        neg = X_int < 0
        off = a.floor()
        p = P[..., : space.n_int]
        z = (p >= base_int).to(X_un.dtype)
        x_new = (-1) ** neg.to(off) * (off + z)
        X_exp[..., idx] = torch.minimum(torch.maximum(x_new, space.integer_bounds[0]), space.integer_bounds[1])
    pos = space.n_int
    for j, (s, c) in enumerate(zip(space.cat_starts, space.cat_cards)):
        cum = P[..., pos:pos + c].cumsum(dim=-1)
        m = cum > base_cat[..., j:j + 1]
        cat = ((m.long().cumsum(dim=-1) == 1).long()).argmax(dim=-1)
        X_exp[..., s:s + c] = torch.nn.functional.one_hot(cat, num_classes=c).to(X_un)
        pos += c
    return X_exp


def one_hot_to_numeric(space: PRSpace, X: Tensor) -> Tensor:
    """input.py:69-88: argmax of every one-hot block -> one numeric code column."""
    if space.n_cat == 0:
        return X
    Xn = X[..., : space.numeric_dim].clone()
    col = space.cat_starts[0]
    for s, c in zip(space.cat_starts, space.cat_cards):
        Xn[..., col] = X[..., s:s + c].argmax(dim=-1)
        col += 1
    return Xn


# ----------------------------------------------------------------------------------------------------------------------
# MC-PR value + gradient  (probabilistic_reparameterization.py:402-648)
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class MCResult:
    value: Tensor  # [b]
    grad: Optional[Tensor]  # [b, 1, d]
    acq_values: Tensor  # [b, mc]
    tilde_x: Tensor  # [b, mc, 1, d]   normalised, one-hot layout
    z: Tensor  # [b, mc, 1, n_disc]   rounding component used by the estimator
    prob: Tensor  # [b, 1, n_disc]
    ma_counter: float  # state AFTER the call
    ma_hidden: float


def mc_pr(
    space: PRSpace,
    X: Tensor,
    acq: Callable[[Tensor], Tensor],
    base_int: Optional[Tensor],
    base_cat: Optional[Tensor],
    grad_estimator: str = "reinforce",
    ma_counter: float = 0.0,
    ma_hidden: float = 0.0,
    apply_numeric: bool = False,
    grad_output: Optional[Tensor] = None,
    with_grad: bool = True,
) -> MCResult:
    """One forward (+ backward) of ``_MCProbabilisticReparameterization`` on ``X [b, 1, d]``.

    * 426-443  only the continuous columns carry an autograd graph
    * 451      tilde_x = normalize(round(unnormalize(X)))
    * 454-465  z: ints ``(tilde - x > 0) | (x == 1)`` in NORMALISED space; cats: the one-hot itself
    * 470-475  prob = rounding_prob(unnormalize(X))
    * 486-498  base AF over all MC samples, mean over MC
    * 499-505  EMA baseline state update (every forward): counter += 1; hidden -= 0.3 (hidden - mean(all))
    * 598-611  grads = mean_i[(a_i - baseline)(z_i - p)] / tau ; baseline from the PRE-update state
    * 613-633  discrete columns: grad_output * grads ; continuous columns: autograd of the MC mean
    """
    X = X.detach().to(torch.float64)
    b, q, d = X.shape
    assert q == 1
    cont = torch.as_tensor(space.cont_indices, dtype=torch.long)
    if len(space.cont_indices) > 0:
        cont_input = X[..., cont].detach().requires_grad_(with_grad)
        cols, ci = [], 0
        for col in range(d):
            if ci < len(space.cont_indices) and col == space.cont_indices[ci]:
                cols.append(cont_input[..., ci])
                ci += 1
            else:
                cols.append(X[..., col])
        Xg = torch.stack(cols, dim=-1)
    else:
        cont_input, Xg = None, X
    tilde = normalize(space, mc_round(space, unnormalize(space, Xg.unsqueeze(-3)), base_int, base_cat))
    rc = tilde.detach().clone()
    if space.n_int > 0:
        idx = torch.as_tensor(space.integer_indices)
        xin = Xg.detach()[..., idx].unsqueeze(-3)
        rc[..., idx] = ((tilde.detach()[..., idx] - xin > 0) | (xin == 1)).to(tilde)
    z = rc[..., torch.as_tensor(space.discrete_indices, dtype=torch.long)]
    prob = rounding_prob(space, unnormalize(space, Xg.detach()))
    tx = one_hot_to_numeric(space, tilde) if apply_numeric else tilde
    acq_values = acq(tx)  # [b, mc]
    mean_acq = acq_values.mean(dim=-1)
    # baseline (pre-update state), then state update
    if grad_estimator == "reinforce":
        baseline = 0.0
    elif grad_estimator == "reinforce_ma":
        baseline = 0.0 if ma_counter == 0 else ma_hidden / (1.0 - 0.7 ** ma_counter)
    else:
        raise NotImplementedError(grad_estimator)  # arm / u2g are dead code in the reference
    new_counter = ma_counter + 1
    new_hidden = ma_hidden - (ma_hidden - float(acq_values.detach().mean())) * (1 - 0.7)
    grad = None
    if with_grad:
        if grad_output is None:
            grad_output = torch.ones(b, dtype=torch.float64)
        av = acq_values.detach().view(b, -1, 1, 1)
        sample_level = (av - baseline) * (z - prob.unsqueeze(-3))
        g = (sample_level / space.tau).mean(dim=-3)  # [b, 1, n_disc]
        grad = grad_output.view(b, 1, 1).expand(b, 1, d).clone()
        grad[..., torch.as_tensor(space.discrete_indices, dtype=torch.long)] *= g
        if cont_input is not None:
            auto = torch.autograd.grad(mean_acq, cont_input, grad_outputs=grad_output)[0]
            grad[..., cont] = auto
    return MCResult(mean_acq.detach(), grad, acq_values.detach(), tilde.detach(), z, prob, new_counter, new_hidden)


# ----------------------------------------------------------------------------------------------------------------------
# Final discrete draw per restart  (probabilistic_reparameterization.py:198-233)
# ----------------------------------------------------------------------------------------------------------------------
def sample_candidates(space: PRSpace, X: Tensor, u: Optional[Tensor] = None) -> Tensor:
    """``X [b, 1, d]`` normalised -> one sampled configuration per restart, normalised.

    * 199-203  un-normalise, rounding probabilities of the UN-normalised point
    * 204-209  integer i: ``floor(x_un) + Bernoulli(p)`` -- floor of the SIGNED value (unlike the MC transform's floor|x|)
    * 210-215  clamp to integer_bounds
    * 217-229  categorical: ``one_hot(Categorical(probs=p).sample())`` written at the one-hot block
    * 231-233  normalise
    ``u is None``: torch's global generator, consumed in the reference's order (one Bernoulli call per integer parameter,
    then one Categorical call per block) -- pinned bit for bit on the golden draws.  ``u [b, n_int + n_cat]``: explicit
    uniforms with ``z = [u < p]`` (torch.bernoulli's rule) and the inverse CDF ``first index with cumsum(p) > u`` (last
    index if none) -- the definition the device kernel ``prb_sample_candidates`` implements for its Philox stream."""
    X = X.detach().to(torch.float64)
    x_un = unnormalize(space, X).clone()
    prob = rounding_prob(space, x_un)
    k = 0
    for i in space.integer_indices:
        p = prob[..., k]
        z = torch.distributions.Bernoulli(probs=p).sample() if u is None else (u[:, k].view_as(p) < p).to(p)
        x_un[..., i] = x_un[..., i].floor() + z
        k += 1
    if space.n_int > 0:
        idx = torch.as_tensor(space.integer_indices)
        x_un[..., idx] = torch.minimum(torch.maximum(x_un[..., idx], space.integer_bounds[0]), space.integer_bounds[1])
    for j, (s, c) in enumerate(zip(space.cat_starts, space.cat_cards)):
        p = prob[..., k:k + c]
        if u is None:
            cat = torch.distributions.Categorical(probs=p).sample()
        else:
            m = p.cumsum(dim=-1) > u[:, space.n_int + j].view(-1, 1, 1)
            first = ((m.long().cumsum(dim=-1) == 1) & m).long().argmax(dim=-1)
            cat = torch.where(m.any(dim=-1), first, torch.full_like(first, c - 1))
        x_un[..., s:s + c] = torch.nn.functional.one_hot(cat, num_classes=c).to(x_un)
        k += c
    return normalize(space, x_un)


# ----------------------------------------------------------------------------------------------------------------------
# Analytic PR  (input.py:273-383, 410-488; probabilistic_reparameterization.py:285-320)
# ----------------------------------------------------------------------------------------------------------------------
def all_discrete_options(space: PRSpace) -> Tensor:
    """Cartesian product of every discrete configuration, categoricals one-hot (input.py:345-383).
    Continuous columns are zero placeholders.  Integer values are UN-normalised."""
    opts = [torch.zeros(1, dtype=torch.long) for _ in range(space.dim - space.n_disc)]
    for i in range(space.n_int):
        opts.append(torch.arange(int(space.integer_bounds[0, i]), int(space.integer_bounds[1, i]) + 1, dtype=torch.long))
    for c in space.cat_cards:
        opts.append(torch.arange(c, dtype=torch.long))
    table = torch.cartesian_prod(*opts)
    if table.ndim == 1:
        table = table.unsqueeze(-1)
    if space.n_cat > 0:
        oh = [torch.nn.functional.one_hot(table[..., k], num_classes=c)
              for k, c in zip((space.categorical_features or {}).keys(), space.cat_cards)]
        table = torch.cat([table[..., : -space.n_cat], *oh], dim=-1)
    return table.to(torch.float64)


def analytic_probs(space: PRSpace, X_un: Tensor, table: Tensor) -> Tensor:
    """P(z | theta) per configuration (input.py:410-464): integer j contributes p if z == floor|x|+1, 1-p if
    z == floor|x|, else 0 (no clamping logic); categorical contributes its softmax entry.  Differentiable.
    X_un [b, 1, d] -> [b, n_discrete]."""
    b = X_un.shape[0]
    nd = table.shape[0]
    probs = torch.ones(b, nd, dtype=X_un.dtype)
    x = X_un[:, 0, :]  # [b, d]
    for k, col in enumerate(space.integer_indices):
        a = x[:, col].abs()
        off = a.floor()
        p = torch.sigmoid((a - off - 0.5) / space.tau)  # [b]
        diff = (off + 1).unsqueeze(-1) - table[:, col].unsqueeze(0)  # [b, nd]
        up = (diff == 0).to(x.dtype)
        down = (diff == 1).to(x.dtype)
        probs = probs * (up * p.unsqueeze(-1) + down * (1 - p.unsqueeze(-1)))
    for s, c in zip(space.cat_starts, space.cat_cards):
        sm = torch.softmax((x[:, s:s + c] - 0.5) / space.tau, dim=-1)  # [b, c]
        sel = table[:, s:s + c]  # [nd, c] one-hot
        probs = probs * (sm.unsqueeze(1) * sel.unsqueeze(0)).sum(-1)
    return probs


def analytic_pr(space: PRSpace, X: Tensor, acq: Callable[[Tensor], Tensor], apply_numeric: bool = False,
                with_grad: bool = True) -> Tuple[Tensor, Optional[Tensor], Tensor, Tensor]:
    """sum_z P(z|theta) acq(x_c, z), autograd end to end (probabilistic_reparameterization.py:285-320).
    Returns (value [b], grad [b,1,d] or None, acq_values [b, nd], probs [b, nd])."""
    X = X.detach().to(torch.float64).clone().requires_grad_(with_grad)
    table = all_discrete_options(space)
    nd = table.shape[0]
    X_un = unnormalize(space, X)
    # transform (input.py:466-488): paste every configuration next to the continuous columns, then normalise
    b = X.shape[0]
    Xe = X_un.unsqueeze(-3).expand(b, nd, 1, space.dim)
    n_disc = space.n_disc
    X_all = torch.cat([Xe[..., : space.dim - n_disc], table.view(1, nd, 1, space.dim).expand(b, nd, 1, space.dim)[..., space.dim - n_disc:]], dim=-1)
    X_all = normalize(space, X_all)
    if apply_numeric:
        X_all = one_hot_to_numeric(space, X_all)
    probs = analytic_probs(space, X_un, table)
    acq_values = acq(X_all)
    value = (acq_values * probs).sum(dim=-1)
    grad = None
    if with_grad:
        grad = torch.autograd.grad(value.sum(), X)[0]
    return value.detach(), grad, acq_values.detach(), probs.detach()


# ----------------------------------------------------------------------------------------------------------------------
# Multi-start Adam (gen.py:304-343) and restart selection (initializers.py:462-606), stochastic=True path only
# ----------------------------------------------------------------------------------------------------------------------
def adam_multistart(value_and_grad: Callable[[Tensor], Tuple[Tensor, Tensor]], value_only: Callable[[Tensor], Tensor],
                    ics: Tensor, lower: Tensor, upper: Tensor, maxiter: int = 200, lr: float = 0.025):
    """clamp -> (value, grad of -sum) -> torch Adam(lr, betas (0.9, 0.999), eps 1e-8) x maxiter -> clamp -> value.
    ``value_and_grad(X[b,1,d]) -> (values [b], d values.sum() / dX)``."""
    clamp = lambda t: torch.max(torch.min(t, upper), lower)
    theta = clamp(ics.detach().clone()).requires_grad_(True)
    opt = torch.optim.Adam([theta], lr=lr)
    for _ in range(maxiter):
        with torch.no_grad():
            Xc = clamp(theta)
        _, g = value_and_grad(Xc)
        opt.zero_grad()
        theta.grad = -g
        opt.step()
    theta = clamp(theta.detach())
    return theta, value_only(theta)


def boltzmann_weights_nonneg(Y: Tensor, eta: float = 1.0, alpha: float = 1e-4):
    """initialize_q_batch_nonneg (initializers.py:537-606): keep Y >= alpha*max, weights exp(eta (Y/max - 1)).
    Returns (candidate index tensor, weights) -- the multinomial draw itself uses torch's global RNG."""
    max_val, max_idx = torch.max(Y, dim=0)
    pos = Y >= alpha * max_val
    idx = torch.nonzero(pos).squeeze(-1)
    alphas = Y[pos]
    weights = torch.exp(eta * (alphas / max_val - 1))
    return idx, weights, int(max_idx)
