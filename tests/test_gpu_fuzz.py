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


def _small_kernel_case(seed):
    """A random problem reshaped to the single-kernel step's domain: no categoricals, mc a multiple of 32, n + 1 <= 128."""
    c = random_case(seed)
    tries = 0
    while (c["categorical_features"] or len(c["integer_indices"]) == 0) and tries < 50:
        tries += 1
        c = random_case(seed + 1000 * tries)
    g = torch.Generator().manual_seed(seed)
    c = dict(c, mc=32 * int(torch.randint(1, 5, (1,), generator=g)), n_train=int(torch.randint(3, 128, (1,), generator=g)),
             b=int(torch.randint(1, 8, (1,), generator=g)), seed=seed)
    return c


@pytest.mark.parametrize("seed", list(range(500, 524)))
def test_random_problem_single_kernel_step_shapes(seed):
    """Shapes that run as ONE kernel per call (prb_small.cu): the separable variant when every integer is binary, the generic
    single-term variant otherwise (ordinals of any range, negative bounds).  Forward + backward and a few fused Adam steps
    against the oracle."""
    c = _small_kernel_case(seed)
    g = build_problem(c)
    space, oacq = oracle_space(g), oracle_acq(g)
    theta, b = g["theta"], c["b"]
    go = torch.linspace(0.7, 1.3, b, dtype=torch.float64)
    ref = P.mc_pr(space, theta, oacq, g["base_samples"], g["base_samples_categorical"],
                  grad_estimator=c["grad_estimator"], ma_counter=1.0, ma_hidden=0.02, apply_numeric=c["apply_numeric"],
                  grad_output=go)
    pr = product_mc_pr(g, DEV)
    pr._ma_state.copy_(torch.tensor([1.0, 0.02], dtype=torch.float64))
    X = theta.to(DEV).requires_grad_(True)
    val = pr(X)
    grad = torch.autograd.grad(val, X, grad_outputs=go.to(DEV))[0]
    scale = max(1.0, float(ref.value.abs().max()))
    ok, w = close(val, ref.value, 1e-9, 1e-12 * scale)
    assert ok, f"value {w:.3g}x tolerance, case {c}"
    ok, w = close(grad, ref.grad, 1e-9, 1e-10 * scale)
    assert ok, f"gradient {w:.3g}x tolerance, case {c}"
    ok, w = close(pr._ma_state, torch.tensor([ref.ma_counter, ref.ma_hidden], dtype=torch.float64), 1e-9, 1e-13)
    assert ok, f"EMA state {w:.3g}x tolerance"
    # fused multi-start Adam (persistent kernel when restarts x mc / 32 <= #SMs) against the oracle's loop
    import bo_pr_b200 as B
    d = c["dim"]
    lo, hi = torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)
    pr2 = product_mc_pr(g, DEV)
    cand, v2 = B.gen_candidates_torch(theta.to(DEV), pr2, lo.to(DEV), hi.to(DEV), options={"maxiter": 5})
    st = {"c": 0.0, "h": 0.0}

    def vg(Xc, with_grad=True):
        r = P.mc_pr(space, Xc, oacq, g["base_samples"], g["base_samples_categorical"], grad_estimator=c["grad_estimator"],
                    ma_counter=st["c"], ma_hidden=st["h"], apply_numeric=c["apply_numeric"], with_grad=with_grad)
        st["c"], st["h"] = r.ma_counter, r.ma_hidden
        return (r.value, r.grad) if with_grad else r.value

    theta_o, val_o = P.adam_multistart(vg, lambda Xc: vg(Xc, False), theta, lo, hi, maxiter=5)
    assert torch.allclose(cand.cpu(), theta_o, rtol=0, atol=1e-8), f"{(cand.cpu() - theta_o).abs().max()} case {c}"
    assert torch.allclose(v2.cpu(), val_o, rtol=1e-7, atol=1e-11), f"case {c}"
