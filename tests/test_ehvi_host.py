"""CPU: the box decomposition behind EHVI (bo_pr_b200/multi_objective.py) and the oracle's EHVI restatement
(oracle/gp_oracle.py) -- against each other, against an independent grid decomposition, and against a Monte Carlo
estimate of the expected hypervolume improvement.  No device calls."""
import math

import pytest
import torch

from bo_pr_b200.multi_objective import FastNondominatedPartitioning, is_non_dominated
from oracle import gp_oracle as G


class _FixedPosterior(torch.nn.Module):
    """A 'model list' with a prescribed posterior, to exercise OracleEHVI on its own."""

    def __init__(self, mean, var):
        super().__init__()
        self.mean, self.var = mean, var

    def posterior(self, X):
        return self.mean.expand(*X.shape[:-2], -1), self.var.expand(*X.shape[:-2], -1)


@pytest.mark.parametrize("m,seed", [(2, 0), (2, 1), (3, 2), (3, 3), (4, 4)])
def test_partitioning_tiles_the_nondominated_region(m, seed):
    g = torch.Generator().manual_seed(seed)
    Y = torch.rand(14, m, generator=g, dtype=torch.float64)
    Y[3] = Y[2]  # a duplicate and a point below the reference point
    Y[5, 0] = 0.05
    ref = torch.full((m,), 0.1, dtype=torch.float64)
    part = FastNondominatedPartitioning(ref, Y)
    cb = part.get_hypercell_bounds()
    assert cb.shape[0] == 2 and cb.shape[2] == m and bool((cb[0] < cb[1]).all())
    P = part.pareto_Y
    assert bool(is_non_dominated(P).all()) and bool((P > ref).all())
    X = 0.1 + 1.1 * torch.rand(20000, m, generator=g, dtype=torch.float64)
    cover = ((X.unsqueeze(1) >= cb[0]) & (X.unsqueeze(1) < cb[1])).all(-1).sum(-1)
    dominated = (P.unsqueeze(0) >= X.unsqueeze(1)).all(-1).any(-1)
    assert bool(((cover == 1) == ~dominated).all()) and int(cover.max()) == 1
    # hypervolume: the dominated part of the box [ref, 1.2]^m by Monte Carlo
    hv = float(part.compute_hypervolume())
    assert abs(hv - float(dominated.double().mean()) * 1.1 ** m) < 0.02


@pytest.mark.parametrize("m,seed", [(2, 10), (3, 11)])
def test_oracle_ehvi_is_decomposition_independent_and_matches_monte_carlo(m, seed):
    g = torch.Generator().manual_seed(seed)
    Y = torch.rand(9, m, generator=g, dtype=torch.float64)
    ref = torch.zeros(m, dtype=torch.float64)
    part = FastNondominatedPartitioning(ref, Y)
    mean = 0.3 + 0.6 * torch.rand(m, generator=g, dtype=torch.float64)
    var = 0.01 + 0.05 * torch.rand(m, generator=g, dtype=torch.float64)
    model = _FixedPosterior(mean, var)
    X = torch.zeros(1, 1, 2, dtype=torch.float64)
    a_sweep = G.OracleEHVI(model, ref, part.get_hypercell_bounds())(X)
    a_grid = G.OracleEHVI(model, ref, G.grid_nondominated_cells(part.pareto_Y, ref))(X)
    assert torch.allclose(a_sweep, a_grid, rtol=1e-12, atol=1e-15)
    # Monte Carlo: E[HV(P + y) - HV(P)] = E[volume of the part of [ref, y] that no Pareto point dominates]
    n = 200000
    ys = mean + var.sqrt() * torch.randn(n, m, generator=g, dtype=torch.float64)
    cb = part.get_hypercell_bounds()
    lo, hi = cb[0], cb[1].clamp_max(1e10)
    side = (torch.minimum(ys.unsqueeze(1), hi) - lo).clamp_min(0.0)  # overlap of [lo, min(y, hi)] per cell and objective
    mc = side.prod(-1).sum(-1)
    est, se = float(mc.mean()), float(mc.std()) / math.sqrt(n)
    assert abs(float(a_sweep) - est) < 5 * se + 1e-6, (float(a_sweep), est, se)
