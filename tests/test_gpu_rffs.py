"""GPU: Thompson sampling over a random-Fourier-feature posterior sample (label ``pr__ts``: rffs.py:24-133,
experiment_utils.py:481-492).  The CUDA evaluation of the sample, PR over it (MC and analytic), and the fused multi-start
Adam are compared with the CPU oracle (`oracle/rff_oracle.py` under `oracle/pr_oracle.py`)."""
import math

import pytest
import torch

import bo_pr_b200 as B
from bo_pr_b200.kernels import MaternKernel, ScaleKernel
from bo_pr_b200.rffs import RandomFourierFeatures, RFFSampleModel, get_gp_samples
from oracle import gp_oracle as G
from oracle import pr_oracle as P
from oracle.rff_oracle import OracleRFFSample
from tests.golden_util import oracle_space
from tests.product_util import close, product_analytic_pr, product_mc_pr
from tests.random_cases import build_problem, random_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sample_model(d_k, m, seed):
    """A random RFF sample over d_k inputs: (device model, CPU oracle of the same function)."""
    torch.manual_seed(seed)
    base = MaternKernel(nu=2.5, ard_num_dims=d_k)
    base.lengthscale = (0.3 + torch.rand(1, d_k, dtype=torch.float64))
    kernel = ScaleKernel(base)
    kernel.outputscale = torch.tensor(1.3, dtype=torch.float64)
    basis = RandomFourierFeatures(kernel, input_dim=d_k, num_rff_features=m, device=torch.device(DEV))
    w = torch.randn(m, dtype=torch.float64)
    model = RFFSampleModel(basis, w.to(DEV))
    return model, OracleRFFSample(basis.weights, basis.bias, w, basis.scale)


@pytest.mark.parametrize("d_k,m,n", [(1, 7, 5), (13, 512, 300), (32, 100, 33)])
def test_sample_path_and_gradient_vs_oracle(d_k, m, n):
    model, oracle = _sample_model(d_k, m, seed=d_k)
    X = torch.rand(n, 1, d_k, dtype=torch.float64)
    Xo = X.clone().requires_grad_(True)
    fo, var_o = oracle.posterior(Xo)
    go = torch.linspace(0.5, 1.5, n, dtype=torch.float64)
    grad_o = torch.autograd.grad(fo, Xo, grad_outputs=go)[0]
    post = model.posterior(X.to(DEV))
    assert post.mean.shape == (n, 1, 1)
    assert torch.allclose(post.mean.reshape(-1).cpu(), fo.detach(), rtol=1e-11, atol=1e-12)
    assert bool((post.variance == 0).all())
    acq = B.PosteriorMean(model)
    Xd = X.to(DEV).requires_grad_(True)
    val = acq(Xd)
    grad = torch.autograd.grad(val, Xd, grad_outputs=go.to(DEV))[0]
    assert torch.allclose(val.cpu(), fo.detach(), rtol=1e-11, atol=1e-12)
    assert torch.allclose(grad.cpu(), grad_o, rtol=1e-10, atol=1e-11)
    with pytest.raises(B._capi.PrbError):  # EI / UCB over a deterministic sample are rejected by the library
        B.UpperConfidenceBound(model, beta=2.0)(X.to(DEV))


@pytest.mark.parametrize("seed", list(range(300, 316)))
def test_mc_pr_over_rff_sample_vs_oracle(seed):
    c = dict(random_case(seed), acq="pm")
    g = build_problem(c)
    d_k = g["train_X_model"].shape[-1]
    model, oracle = _sample_model(d_k, 64 + 8 * (seed % 5), seed)
    space, oacq = oracle_space(g), G.OraclePosteriorMean(oracle)
    theta, b = g["theta"], c["b"]
    go = torch.linspace(0.7, 1.3, b, dtype=torch.float64)
    ref = P.mc_pr(space, theta, oacq, g["base_samples"], g["base_samples_categorical"],
                  grad_estimator=c["grad_estimator"], ma_counter=3.0, ma_hidden=-0.2, apply_numeric=c["apply_numeric"],
                  grad_output=go)
    pr = product_mc_pr(g, DEV, acq=B.PosteriorMean(model))
    pr._ma_state.copy_(torch.tensor([3.0, -0.2], dtype=torch.float64))
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


def _n_options(c):
    n = 1
    for lo, hi in zip(*c["integer_bounds"]):
        n *= int(hi - lo) + 1
    for card in (c["categorical_features"] or {}).values():
        n *= card
    return n


@pytest.mark.parametrize("seed", [s for s in range(400, 460) if _n_options(random_case(s)) <= 300][:6])
def test_analytic_pr_over_rff_sample_vs_oracle(seed):
    c = dict(random_case(seed), acq="pm")
    g = build_problem(c)
    d_k = g["train_X_model"].shape[-1]
    model, oracle = _sample_model(d_k, 96, seed)
    space, oacq = oracle_space(g), G.OraclePosteriorMean(oracle)
    v_ref, g_ref, _, _ = P.analytic_pr(space, g["theta"], oacq, apply_numeric=c["apply_numeric"])
    pr = product_analytic_pr(g, DEV, acq=B.PosteriorMean(model))
    X = g["theta"].to(DEV).requires_grad_(True)
    val = pr(X)
    grad = torch.autograd.grad(val.sum(), X)[0]
    scale = max(1.0, float(v_ref.abs().max()))
    ok, w = close(val, v_ref, 1e-9, 1e-12 * scale)
    assert ok, f"value {w:.3g}x tolerance, case {c}"
    ok, w = close(grad, g_ref, 1e-9, 1e-10 * scale)
    assert ok, f"gradient {w:.3g}x tolerance, case {c}"


def test_fused_adam_over_rff_sample_matches_host_loop_and_oracle():
    c = dict(random_case(7), acq="pm", grad_estimator="reinforce_ma", b=4, mc=64)
    g = build_problem(c)
    d = c["dim"]
    model, oracle = _sample_model(g["train_X_model"].shape[-1], 128, 5)
    bounds = torch.stack([torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)]).to(DEV)
    ics = g["theta"].to(DEV)
    steps = 10
    pr_f = product_mc_pr(g, DEV, acq=B.PosteriorMean(model))
    cand_f, val_f = B.gen_candidates_torch(ics, pr_f, bounds[0], bounds[1], options={"maxiter": steps})
    pr_h = product_mc_pr(g, DEV, acq=B.PosteriorMean(model))
    cand_h, val_h = B.gen_candidates_torch(ics, pr_h, bounds[0], bounds[1], options={"maxiter": steps, "fused": False})
    assert torch.allclose(cand_f, cand_h, rtol=0, atol=1e-10), (cand_f - cand_h).abs().max()
    assert torch.allclose(val_f, val_h, rtol=1e-8, atol=1e-12)
    space, oacq = oracle_space(g), G.OraclePosteriorMean(oracle)
    st = {"c": 0.0, "h": 0.0}

    def vg(X, with_grad=True):
        r = P.mc_pr(space, X, oacq, g["base_samples"], g["base_samples_categorical"], grad_estimator="reinforce_ma",
                    ma_counter=st["c"], ma_hidden=st["h"], apply_numeric=c["apply_numeric"], with_grad=with_grad)
        st["c"], st["h"] = r.ma_counter, r.ma_hidden
        return (r.value, r.grad) if with_grad else r.value

    theta_o, val_o = P.adam_multistart(vg, lambda X: vg(X, False), g["theta"], torch.zeros(d, dtype=torch.float64),
                                       torch.ones(d, dtype=torch.float64), maxiter=steps)
    assert torch.allclose(cand_f.cpu(), theta_o, rtol=0, atol=1e-8), (cand_f.cpu() - theta_o).abs().max()
    assert torch.allclose(val_f.cpu(), val_o, rtol=1e-7, atol=1e-11)


def test_get_gp_samples_draws_functions_that_explain_the_data():
    """End to end: FixedNoiseGP with a Scale(Matern) kernel -> get_gp_samples -> PosteriorMean; at low noise every posterior
    sample passes close to the training targets, and two draws differ away from the data."""
    torch.manual_seed(11)
    n, d = 40, 3
    X = torch.rand(n, d, dtype=torch.float64, device=DEV)
    y = torch.sin(3 * X[:, 0]) + X[:, 1] ** 2 - 0.5 * X[:, 2]
    y = (y - y.mean()) / y.std()
    base = MaternKernel(nu=2.5, ard_num_dims=d)
    base.lengthscale = torch.tensor([[0.6, 0.8, 1.0]], dtype=torch.float64)
    covar = ScaleKernel(base)
    covar.outputscale = torch.tensor(1.0, dtype=torch.float64)
    gp = B.FixedNoiseGP(X, y.unsqueeze(-1), torch.full((n, 1), 1e-4, dtype=torch.float64, device=DEV), covar_module=covar)
    s1 = get_gp_samples(gp, num_outputs=1, n_samples=1, num_rff_features=512)
    s2 = get_gp_samples(gp, num_outputs=1, n_samples=1, num_rff_features=512)
    f1 = s1.posterior(X.unsqueeze(-2)).mean.reshape(-1)
    assert float((f1 - y).abs().max()) < 0.15, float((f1 - y).abs().max())
    Xt = torch.rand(200, 1, d, dtype=torch.float64, device=DEV) * 3 + 2  # far from the data: prior-like variation
    assert float((s1.posterior(Xt).mean - s2.posterior(Xt).mean).abs().mean()) > 0.1
    acq = B.PosteriorMean(s1)
    assert acq(X.unsqueeze(-2)).shape == (n,)
    with pytest.raises(NotImplementedError):
        get_gp_samples(gp, num_outputs=1, n_samples=4)
