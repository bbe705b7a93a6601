"""Host-side logic of the optimisation drivers (CPU, toy acquisition functions): restart selection heuristics, stopping
criterion, fixed-feature wrapper, the host-driven Adam loop against the oracle's restatement of gen.py:304-343."""
import math
import warnings

import pytest
import torch

from bo_pr_b200.gen import (ExpMAStoppingCriterion, FixedFeatureAcquisitionFunction, columnwise_clamp,
                            gen_candidates_scipy, gen_candidates_torch, get_best_candidates)
from bo_pr_b200.initializers import (BadInitialCandidatesWarning, draw_sobol_samples, fix_features,
                                     gen_batch_initial_conditions, initialize_q_batch, initialize_q_batch_nonneg)
from bo_pr_b200.optimize import optimize_acqf, optimize_acqf_mixed
from oracle import pr_oracle as P


class Quad(torch.nn.Module):
    """acq(X[b, 1, d]) = 1 - ||x - c||^2 (max 1 at c)."""

    def __init__(self, c):
        super().__init__()
        self.c = torch.as_tensor(c, dtype=torch.float64)
        self.X_pending = None
        self.calls = 0

    def set_X_pending(self, X_pending=None):
        self.X_pending = X_pending

    def forward(self, X):
        self.calls += 1
        return 1.0 - ((X - self.c) ** 2).sum(dim=(-1, -2))


def test_initialize_q_batch_nonneg_matches_oracle_weights_and_keeps_argmax():
    g = torch.Generator().manual_seed(0)
    X = torch.rand(200, 1, 4, generator=g, dtype=torch.float64)
    Y = torch.rand(200, generator=g, dtype=torch.float64) ** 8
    idx, weights, max_idx = P.boltzmann_weights_nonneg(Y)
    torch.manual_seed(5)
    expect = idx[torch.multinomial(weights, 10)]
    if max_idx not in expect:
        expect[-1] = max_idx
    torch.manual_seed(5)
    got = initialize_q_batch_nonneg(X, Y, 10)
    assert torch.equal(got, X[expect])
    assert any(torch.equal(r, X[max_idx]) for r in got)
    assert torch.equal(initialize_q_batch_nonneg(X, Y, 200), X)
    with pytest.raises(RuntimeError):
        initialize_q_batch_nonneg(X, Y, 201)


def test_initialize_q_batch_nonneg_degenerate_inputs():
    X = torch.rand(50, 1, 3, dtype=torch.float64)
    with warnings.catch_warnings(record=True) as ws:
        warnings.simplefilter("always")
        out = initialize_q_batch_nonneg(X, torch.zeros(50, dtype=torch.float64), 5)
    assert out.shape == (5, 1, 3) and any(issubclass(w.category, BadInitialCandidatesWarning) for w in ws)
    Y = torch.zeros(50, dtype=torch.float64)
    Y[[3, 7]] = 1.0  # fewer positive points than requested: all of them are kept
    out = initialize_q_batch_nonneg(X, Y, 5)
    assert out.shape == (5, 1, 3) and any(torch.equal(r, X[3]) for r in out) and any(torch.equal(r, X[7]) for r in out)


def test_initialize_q_batch_zscore_weights():
    g = torch.Generator().manual_seed(1)
    X = torch.rand(100, 1, 2, generator=g, dtype=torch.float64)
    Y = torch.randn(100, generator=g, dtype=torch.float64)
    torch.manual_seed(3)
    w = torch.exp((Y - Y.mean()) / Y.std())
    expect = torch.multinomial(w, 8)
    if Y.argmax() not in expect:
        expect[-1] = Y.argmax()
    torch.manual_seed(3)
    assert torch.equal(initialize_q_batch(X, Y, 8), X[expect])
    with warnings.catch_warnings(record=True) as ws:
        warnings.simplefilter("always")
        initialize_q_batch(X, torch.ones(100, dtype=torch.float64), 8)
    assert any(issubclass(w_.category, BadInitialCandidatesWarning) for w_ in ws)


def test_stopping_criterion_fixed_iterations_and_relative_tolerance():
    s = ExpMAStoppingCriterion(maxiter=7, rel_tol=float("-inf"))
    assert [s.evaluate(torch.tensor(1.0 / (i + 1))) for i in range(7)] == [False] * 6 + [True]
    s = ExpMAStoppingCriterion(maxiter=1000, rel_tol=1e-3)
    steps = 0
    while not s.evaluate(torch.tensor(1.0 + math.exp(-steps))):
        steps += 1
    assert 10 < steps < 40  # converging sequence stops once the windowed average stalls


def test_fixed_feature_wrapper_and_fix_features():
    acq = Quad([0.1, 0.2, 0.3, 0.4])
    ff = FixedFeatureAcquisitionFunction(acq, d=4, columns=[3, 1], values=[0.4, 0.2])
    Xr = torch.tensor([[[0.1, 0.3]], [[0.5, 0.5]]], dtype=torch.float64)
    full = ff._construct_X_full(Xr)
    assert torch.equal(full, torch.tensor([[[0.1, 0.2, 0.3, 0.4]], [[0.5, 0.2, 0.5, 0.4]]], dtype=torch.float64))
    assert torch.allclose(ff(Xr), acq(full))
    X = torch.rand(5, 1, 4, dtype=torch.float64)
    Xf = fix_features(X, {0: 0.25})
    assert bool((Xf[..., 0] == 0.25).all()) and torch.equal(Xf[..., 1:], X[..., 1:])


def test_gen_candidates_torch_host_loop_matches_oracle_adam():
    acq = Quad([0.3, 0.9, 0.5])
    ics = torch.tensor([[[0.0, 0.0, 0.0]], [[1.0, 1.0, 1.0]], [[0.2, 0.95, 0.4]]], dtype=torch.float64)
    lo, hi = torch.zeros(3, dtype=torch.float64), torch.full((3,), 0.8, dtype=torch.float64)

    def vg(X):
        Xr = X.clone().requires_grad_(True)
        v = acq(Xr)
        return v.detach(), torch.autograd.grad(v.sum(), Xr)[0]

    ref_theta, ref_val = P.adam_multistart(vg, lambda X: acq(X), ics, lo, hi, maxiter=50, lr=0.025)
    opts = {"maxiter": 50}
    cand, vals = gen_candidates_torch(ics, acq, lo, hi, options=opts)
    assert opts["rel_tol"] == float("-inf")  # the reference's side effect on the caller's dict (gen.py:311-313)
    assert torch.allclose(cand, ref_theta, rtol=0, atol=1e-14) and torch.allclose(vals, ref_val, rtol=0, atol=1e-14)
    assert bool((cand <= 0.8).all()) and bool((cand >= 0).all())
    seen = []
    gen_candidates_torch(ics, acq, lo, hi, options={"maxiter": 3}, callback=lambda i, loss, grad, X: seen.append(i))
    assert seen == [1, 2, 3]
    cand_ff, _ = gen_candidates_torch(ics, acq, lo, hi, options={"maxiter": 5}, fixed_features={1: 0.7})
    assert bool((cand_ff[..., 1] == 0.7).all())
    assert torch.equal(get_best_candidates(cand, vals), cand[vals.argmax()])


def test_gen_candidates_scipy_box_constrained_optimum():
    acq = Quad([0.3, 0.9, 0.5])
    ics = torch.tensor([[[0.1, 0.1, 0.1]]], dtype=torch.float64)
    lo, hi = torch.zeros(3, dtype=torch.float64), torch.full((3,), 0.8, dtype=torch.float64)
    cand, vals = gen_candidates_scipy(ics, acq, lo, hi)
    assert torch.allclose(cand, torch.tensor([[[0.3, 0.8, 0.5]]], dtype=torch.float64), atol=1e-6)
    with pytest.raises(NotImplementedError):
        gen_candidates_scipy(ics, acq, lo, hi, inequality_constraints=[(None, None, 0.0)])


def test_optimize_acqf_shapes_screening_chunks_and_argmax():
    acq = Quad([0.3, 0.6])
    bounds = torch.tensor([[0.0, 0.0], [1.0, 1.0]], dtype=torch.float64)
    opts = {"batch_limit": 2, "init_batch_limit": 16, "maxiter": 40, "nonnegative": True, "seed": 0}
    cand, vals = optimize_acqf(acq, bounds, q=1, num_restarts=5, raw_samples=64, options=opts, stochastic=True,
                               return_best_only=False)
    assert cand.shape == (5, 1, 2) and vals.shape == (5,)
    # 64 / 16 screening calls + per batch of 2 restarts: 40 steps + the final evaluation (3 batches)
    assert acq.calls == 4 + 3 * 41
    best, bv = optimize_acqf(acq, bounds, q=1, num_restarts=5, raw_samples=64, options=dict(opts), stochastic=True)
    assert best.shape == (1, 2) and float(bv) > 0.95
    with pytest.raises(ValueError):
        optimize_acqf(acq, bounds[0], q=1, num_restarts=2, raw_samples=8)
    with pytest.raises(ValueError):
        optimize_acqf(acq, bounds, q=1, num_restarts=2)
    ics = gen_batch_initial_conditions(acq, bounds, q=1, num_restarts=3, raw_samples=32, options={"seed": 1})
    assert ics.shape == (3, 1, 2)
    assert draw_sobol_samples(bounds, 8, 1, seed=3).shape == (8, 1, 2)
    assert torch.equal(columnwise_clamp(torch.tensor([[-1.0, 2.0]]), 0.0, 1.0), torch.tensor([[0.0, 1.0]]))


def test_optimize_acqf_mixed_enumerates_fixed_features():
    acq = Quad([0.3, 0.6, 1.0])
    bounds = torch.tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 2.0]], dtype=torch.float64)
    cand, val = optimize_acqf_mixed(acq, bounds, q=1, num_restarts=3, raw_samples=32,
                                    fixed_features_list=[{2: 0.0}, {2: 1.0}, {2: 2.0}],
                                    options={"maxiter": 60, "seed": 0}, stochastic=True)
    assert cand.shape == (1, 3) and float(cand[0, 2]) == 1.0 and float(val) > 0.95
    cand2, val2 = optimize_acqf_mixed(acq, bounds, q=1, num_restarts=2, raw_samples=16,
                                      fixed_features_list=[{2: 0.0}, {2: 1.0}], options={"seed": 0})  # L-BFGS-B
    assert float(cand2[0, 2]) == 1.0 and float(val2) > 0.999
    with pytest.raises(ValueError):
        optimize_acqf_mixed(acq, bounds, q=1, num_restarts=2, fixed_features_list=[])
