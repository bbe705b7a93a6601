"""Differential test over seeded random PR problems: random mixes of continuous / binary / ordinal (incl. negative bounds) /
categorical parameters, both kernel structures of get_kernel, EI / UCB / posterior mean, both estimators, ragged sizes.
CUDA path (reference-API classes -> C ABI) vs the CPU oracle: samples bit-exact, values / gradients to 1e-9."""
import pytest
import torch

from oracle import pr_oracle as P
from tests.golden_util import oracle_acq, oracle_space
from tests.product_util import close, product_mc_pr
from tests.random_cases import build_problem, random_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("seed", list(range(40)))
def test_random_problem(seed):
    c = random_case(seed)
    g = build_problem(c)
    space, oacq = oracle_space(g), oracle_acq(g)
    theta = g["theta"]
    b = c["b"]
    go = torch.linspace(0.7, 1.3, b, dtype=torch.float64)
    ref = P.mc_pr(space, theta, oacq, g["base_samples"], g["base_samples_categorical"],
                  grad_estimator=c["grad_estimator"], ma_counter=2.0, ma_hidden=0.04, apply_numeric=c["apply_numeric"],
                  grad_output=go)
    pr = product_mc_pr(g, DEV)
    pr._ma_state.copy_(torch.tensor([2.0, 0.04], dtype=torch.float64))
    X = theta.to(DEV).requires_grad_(True)
    val = pr(X)
    grad = torch.autograd.grad(val, X, grad_outputs=go.to(DEV))[0]
    tilde = pr.input_transform(theta.to(DEV).unsqueeze(-3))
    assert torch.equal(tilde.cpu(), ref.tilde_x), f"case {c}"
    scale = max(1.0, float(ref.value.abs().max()))
    ok, w = close(val, ref.value, 1e-9, 1e-12 * scale)
    assert ok, f"value {w:.3g}x tolerance, case {c}"
    ok, w = close(grad, ref.grad, 1e-9, 1e-10 * scale)
    assert ok, f"gradient {w:.3g}x tolerance, case {c}"
    ok, w = close(pr._ma_state, torch.tensor([ref.ma_counter, ref.ma_hidden], dtype=torch.float64), 1e-9, 1e-13)
    assert ok, f"EMA state {w:.3g}x tolerance"


def _n_options(c):
    n = 1
    for lo, hi in zip(*c["integer_bounds"]):
        n *= int(hi - lo) + 1
    for card in (c["categorical_features"] or {}).values():
        n *= card
    return n


ANALYTIC_SEEDS = [s for s in range(100, 160) if _n_options(random_case(s)) <= 300][:12]


@pytest.mark.parametrize("seed", ANALYTIC_SEEDS)
def test_random_problem_analytic(seed):
    from tests.product_util import product_analytic_pr
    c = random_case(seed)
    g = build_problem(c)
    space, oacq = oracle_space(g), oracle_acq(g)
    theta = g["theta"]
    theta[:, 0, c["integer_indices"]] = theta[:, 0, c["integer_indices"]].clamp(0.02, 0.98)  # interior: no lost mass
    value, grad, _, probs = P.analytic_pr(space, theta, oacq, apply_numeric=c["apply_numeric"])
    pr = product_analytic_pr(g, DEV)
    X = theta.to(DEV).requires_grad_(True)
    val = pr(X)
    gr = torch.autograd.grad(val.sum(), X)[0]
    scale = max(1.0, float(value.abs().max()))
    ok, w = close(val, value, 1e-9, 1e-12 * scale)
    assert ok, f"analytic value {w:.3g}x tolerance, case {c}"
    ok, w = close(gr, grad, 1e-9, 1e-9 * scale)
    assert ok, f"analytic gradient {w:.3g}x tolerance, case {c}"
