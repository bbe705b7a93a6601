"""Trust-region host logic against the transition rules of ``discrete_mixed_bo/trust_region.py:16-85`` and
the bound computation of ``run_one_replication.py:410-440`` (expected values worked out by hand from those rules)."""
import math

import torch

from bo_pr_b200.trust_region import TurboState, trust_region_bounds, trust_region_center, update_state


def test_state_defaults_and_failure_tolerance():
    s = TurboState(dim=16, batch_size=1, is_constrained=False)
    assert (s.length, s.length_min, s.length_max, s.success_tolerance) == (0.8, 0.5**7, 1.6, 10)
    assert s.failure_tolerance == 16
    assert TurboState(dim=2, batch_size=1, is_constrained=False).failure_tolerance == 4
    assert TurboState(dim=13, batch_size=4, is_constrained=False).failure_tolerance == math.ceil(13 / 4)


def test_unconstrained_shrink_expand_restart():
    s = TurboState(dim=2, batch_size=1, is_constrained=False)  # failure tolerance 4
    # reference quirk kept: against best_value = -inf the threshold is -inf + 1e-3 * inf = NaN, so the very first
    # observation never counts as a success (trust_region.py:40-42)
    update_state(s, torch.tensor([[1.0]], dtype=torch.float64))
    assert s.best_value == 1.0 and s.success_counter == 0 and s.failure_counter == 1
    update_state(s, torch.tensor([[1.0005]], dtype=torch.float64))  # below the 1e-3 relative threshold: a failure, but best_value moves
    assert s.best_value == 1.0005 and s.success_counter == 0 and s.failure_counter == 2
    update_state(s, torch.tensor([[1.1]], dtype=torch.float64))
    assert s.best_value == 1.1 and s.success_counter == 1 and s.failure_counter == 0
    for _ in range(4):
        update_state(s, torch.tensor([[0.0]], dtype=torch.float64))
    assert s.length == 0.4 and s.failure_counter == 0
    y = 2.0
    for _ in range(10):
        y *= 1.1
        update_state(s, torch.tensor([[y]], dtype=torch.float64))
    assert s.length == 0.8 and s.success_counter == 0
    for _ in range(20):
        y *= 1.1
        update_state(s, torch.tensor([[y]], dtype=torch.float64))
    assert s.length == 1.6  # capped at length_max
    assert not s.restart_triggered
    for _ in range(4 * 8):
        update_state(s, torch.tensor([[0.0]], dtype=torch.float64))
    assert s.length == 1.6 / 2**8 and s.restart_triggered


def test_constrained_cases():
    s = TurboState(dim=4, batch_size=1, is_constrained=True)
    update_state(s, torch.tensor([[5.0, -1.0, -0.5]], dtype=torch.float64))  # infeasible; inf - 1e-3 * inf = NaN: not a success (same quirk)
    assert s.constraint_violation == 1.5 and s.failure_counter == 1 and s.best_value == -float("inf")
    update_state(s, torch.tensor([[9.0, -1.0, -0.4999]], dtype=torch.float64))  # 1.4999 is not a 1e-3 relative improvement
    assert s.constraint_violation == 1.4999 and s.success_counter == 0 and s.failure_counter == 2
    update_state(s, torch.tensor([[9.0, -1.0, 0.0]], dtype=torch.float64))  # violation 1.0: success
    assert s.constraint_violation == 1.0 and s.success_counter == 1 and s.failure_counter == 0
    update_state(s, torch.tensor([[0.3, 0.0, 2.0], [7.0, -1.0, 1.0]], dtype=torch.float64))  # first feasible point
    assert s.constraint_violation == 0.0 and s.best_value == 0.3 and s.success_counter == 2
    update_state(s, torch.tensor([[9.0, -1.0, 1.0]], dtype=torch.float64))  # infeasible newcomer after a feasible incumbent
    assert s.best_value == 0.3 and s.failure_counter == 1
    update_state(s, torch.tensor([[0.5, 1.0, 1.0], [8.0, -1.0, 0.0]], dtype=torch.float64))
    assert s.best_value == 0.5 and s.success_counter == 1


def test_bounds_follow_the_incumbent():
    d = 6
    std = torch.stack([torch.zeros(d), torch.ones(d)]).double()
    X = torch.tensor([[0.1, 0.5, 0.9, 1.0, 0.0, 1.0], [0.2, 0.6, 0.95, 0.0, 1.0, 0.0]], dtype=torch.float64).double()
    Y = torch.tensor([[0.0], [1.0]], dtype=torch.float64).double()
    s = TurboState(dim=d, batch_size=1, is_constrained=False, length=0.25)
    b = trust_region_bounds(s, X, Y, std, binary_dims=[3], categorical_start=4)
    assert torch.allclose(b[0], torch.tensor([0.0, 0.35, 0.70, 0.0, 0.0, 0.0]).double())
    assert torch.allclose(b[1], torch.tensor([0.45, 0.85, 1.0, 1.0, 1.0, 1.0]).double())
    Yc = torch.tensor([[5.0, -1.0], [1.0, -0.2]], dtype=torch.float64).double()
    assert torch.equal(trust_region_center(X, Yc, True), X[1])  # nothing feasible: least violation
    Yc = torch.tensor([[5.0, 0.1], [9.0, -0.2]], dtype=torch.float64).double()
    assert torch.equal(trust_region_center(X, Yc, True), X[0])  # best among the feasible
