"""Model lists under MC-PR, CUDA path vs the oracle: ConstrainedExpectedImprovement over independent GPs
(experiment_utils.py:354-392 -- the welded-beam / pressure-vessel style configs) and analytic ExpectedHypervolumeImprovement
(experiment_utils.py:395-402 -- the multi-objective configs, label pr__ehvi)."""
import pytest
import torch

import bo_pr_b200 as B
from oracle import gp_oracle as G
from oracle import pr_oracle as P
from tests.golden_util import oracle_space, oracle_spec
from tests.product_util import close, integer_bounds, product_model
from tests.random_cases import build_problem

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CASES = {
    # welded-beam layout: binary + ordinals + one categorical, NO continuous parameter (score-function gradient only)
    "welded_like": dict(dim=8, integer_indices=[0, 1, 2, 3], integer_bounds=[[0.0, 0.0, 0.0, 0.0], [1.0, 7.0, 12.0, 5.0]],
                        categorical_features={4: 4}, kernel_type="mixed", binary_dims=[0], n_train=45, b=5, mc=64,
                        acq="ei", apply_numeric=False, grad_estimator="reinforce_ma", seed=51, beta=1.0),
    # two continuous parameters: the pathwise gradient flows through every model of the list; two row tiles
    "mixed_cont": dict(dim=6, integer_indices=[2, 3, 4, 5], integer_bounds=[[0.0, 0.0, -1.0, 0.0], [1.0, 1.0, 2.0, 4.0]],
                       categorical_features=None, kernel_type="mixed", binary_dims=[2, 3], n_train=150, b=4, mc=48,
                       acq="ei", apply_numeric=False, grad_estimator="reinforce", seed=52, beta=1.0),
}


def _outputs(c, n_out=3):
    """Problem dicts for the objective + constraints: same inputs, different targets and hyper-parameters."""
    g0 = build_problem(c)
    gs = [g0]
    gen = torch.Generator().manual_seed(c["seed"] + 5)
    X = g0["train_X_model"]
    for m in range(1, n_out):
        w = torch.randn(X.shape[-1], generator=gen, dtype=torch.float64)
        y = torch.cos(2.0 * X @ w + m) + 0.05 * torch.randn(X.shape[0], generator=gen, dtype=torch.float64)
        hp = dict(g0["hp"])
        hp["cont_lengthscale"] = [l * (0.8 + 0.3 * m) for l in hp["cont_lengthscale"]]
        hp["outputscale"] = hp["outputscale"] * (1.0 + 0.25 * m)
        gs.append(dict(g0, train_Y=(y - y.mean()) / y.std(), hp=hp, mean_const=0.1 * m))
    return gs


def _build(name):
    c = CASES[name]
    gs = _outputs(c)
    ogps = [G.OracleGP(g["train_X_model"], g["train_Y"], oracle_spec(g), g["noise"], g["mean_const"]) for g in gs]
    cons = {1: (-0.3, None), 2: (-1.0, 0.8)}
    best_f = float(gs[0]["train_Y"].quantile(0.7))
    oacq = G.OracleConstrainedEI(G.OracleModelList(*ogps), best_f, 0, cons)
    model = B.ModelListGP(*[product_model(g, DEV) for g in gs])
    acq = B.ConstrainedExpectedImprovement(model, best_f=best_f, objective_index=0, constraints=cons)
    pr = B.MCProbabilisticReparameterization(
        acq_function=acq, dim=c["dim"], integer_indices=c["integer_indices"], integer_bounds=integer_bounds(c, DEV),
        categorical_features=c["categorical_features"], mc_samples=c["mc"], grad_estimator=c["grad_estimator"], tau=0.1)
    rnd = pr.input_transform["round"]
    g0 = gs[0]
    if g0["base_samples"] is not None:
        rnd.register_buffer("base_samples", g0["base_samples"].to(DEV).contiguous())
    if g0["base_samples_categorical"] is not None:
        rnd.register_buffer("base_samples_categorical", g0["base_samples_categorical"].to(DEV).contiguous())
    return c, g0, oacq, acq, pr


@pytest.mark.parametrize("name", list(CASES))
def test_constrained_ei_points_vs_oracle(name):
    c, g0, oacq, acq, pr = _build(name)
    gen = torch.Generator().manual_seed(3)
    Xk = g0["train_X_model"]
    idx = torch.randint(0, Xk.shape[0], (40,), generator=gen)
    Xt = Xk[idx].clone()
    cont = g0["cont_dims"]
    if cont:
        Xt[:, cont] = (Xt[:, cont] + 0.2 * torch.randn(40, len(cont), generator=gen, dtype=torch.float64)).clamp(0, 1)
    with torch.no_grad():
        ref = oacq(Xt.unsqueeze(-2))
        out = acq(Xt.to(DEV).unsqueeze(-2))
    ok, worst = close(out, ref, 1e-9, 1e-13)
    assert ok, f"constrained EI at explicit points: {worst:.3g} x tolerance"
    post = acq.model.posterior(Xt.to(DEV).unsqueeze(-2))
    m_ref, v_ref = oacq.model.posterior(Xt.unsqueeze(-2))
    assert close(post.mean.reshape(40, -1), m_ref.reshape(40, -1))[0]
    assert close(post.variance.reshape(40, -1), v_ref.reshape(40, -1))[0]


@pytest.mark.parametrize("name", list(CASES))
def test_mc_pr_over_model_list_vs_oracle(name):
    c, g0, oacq, acq, pr = _build(name)
    space = oracle_space(g0)
    theta = g0["theta"]
    go = torch.linspace(0.6, 1.4, c["b"], dtype=torch.float64)
    pre = (3.0, 0.02) if c["grad_estimator"] == "reinforce_ma" else (0.0, 0.0)
    ref = P.mc_pr(space, theta, oacq, g0["base_samples"], g0["base_samples_categorical"],
                  grad_estimator=c["grad_estimator"], ma_counter=pre[0], ma_hidden=pre[1], grad_output=go)
    pr._ma_state.copy_(torch.tensor(pre, dtype=torch.float64))
    X = theta.to(DEV).requires_grad_(True)
    val = pr(X)
    grad = torch.autograd.grad(val, X, grad_outputs=go.to(DEV))[0]
    ok, w = close(val, ref.value, 1e-9, 1e-13)
    assert ok, f"PR value over the model list: {w:.3g} x tolerance"
    ok, w = close(grad, ref.grad, 1e-9, 1e-12)
    assert ok, f"PR gradient over the model list: {w:.3g} x tolerance"
    ok, w = close(pr._ma_state, torch.tensor([ref.ma_counter, ref.ma_hidden], dtype=torch.float64), 1e-12, 0.0)
    assert ok
    # a few Adam steps through the public optimiser entry (host loop: model lists are not fused) improve the objective
    lo = torch.zeros(c["dim"], dtype=torch.float64, device=DEV)
    hi = torch.ones(c["dim"], dtype=torch.float64, device=DEV)
    with torch.no_grad():
        v0 = pr(theta.to(DEV))
    cand, v1 = B.gen_candidates_torch(theta.to(DEV), pr, lo, hi, options={"maxiter": 10})
    assert cand.shape == theta.shape and bool((cand >= 0).all()) and bool((cand <= 1).all())
    assert float(v1.sum()) >= float(v0.sum()) - 1e-9
    chosen, value = B.finalize_candidates(pr, cand, seed=1)
    assert chosen.shape == (1, c["dim"]) and float(value) >= 0.0


# ---------------------------------------------------------------------------------------------------------------------
# Expected hypervolume improvement (experiment_utils.py:395-402, label pr__ehvi): every model an objective.
# Oracle: the BoTorch formula restated (2^m cross product) over an INDEPENDENT grid decomposition of the non-dominated
# region; product: multi_ehvi_epilogue_kernel over the dimension-sweep decomposition (bo_pr_b200/multi_objective.py).
# ---------------------------------------------------------------------------------------------------------------------
def _build_ehvi(name, n_obj):
    c = CASES[name]
    gs = _outputs(c, n_out=n_obj)
    ogps = [G.OracleGP(g["train_X_model"], g["train_Y"], oracle_spec(g), g["noise"], g["mean_const"]) for g in gs]
    Y = torch.stack([g["train_Y"] for g in gs], dim=-1)
    ref = Y.quantile(0.2, dim=0)
    part = B.FastNondominatedPartitioning(ref_point=ref, Y=Y)
    assert part.pareto_Y.shape[0] >= 2
    oacq = G.OracleEHVI(G.OracleModelList(*ogps), ref, G.grid_nondominated_cells(part.pareto_Y, ref))
    model = B.ModelListGP(*[product_model(g, DEV) for g in gs])
    acq = B.ExpectedHypervolumeImprovement(model, ref_point=ref.tolist(), partitioning=part)
    pr = B.MCProbabilisticReparameterization(
        acq_function=acq, dim=c["dim"], integer_indices=c["integer_indices"], integer_bounds=integer_bounds(c, DEV),
        categorical_features=c["categorical_features"], mc_samples=c["mc"], grad_estimator=c["grad_estimator"], tau=0.1)
    rnd = pr.input_transform["round"]
    g0 = gs[0]
    if g0["base_samples"] is not None:
        rnd.register_buffer("base_samples", g0["base_samples"].to(DEV).contiguous())
    if g0["base_samples_categorical"] is not None:
        rnd.register_buffer("base_samples_categorical", g0["base_samples_categorical"].to(DEV).contiguous())
    return c, g0, oacq, acq, pr


@pytest.mark.parametrize("name,n_obj", [("welded_like", 2), ("mixed_cont", 2), ("mixed_cont", 3)])
def test_ehvi_points_vs_oracle(name, n_obj):
    c, g0, oacq, acq, pr = _build_ehvi(name, n_obj)
    gen = torch.Generator().manual_seed(4)
    Xk = g0["train_X_model"]
    idx = torch.randint(0, Xk.shape[0], (40,), generator=gen)
    Xt = Xk[idx].clone()
    cont = g0["cont_dims"]
    if cont:
        Xt[:, cont] = (Xt[:, cont] + 0.2 * torch.randn(40, len(cont), generator=gen, dtype=torch.float64)).clamp(0, 1)
    with torch.no_grad():
        ref = oacq(Xt.unsqueeze(-2))
        out = acq(Xt.to(DEV).unsqueeze(-2))
    assert float(ref.max()) > 0.0
    ok, worst = close(out, ref, 1e-9, 1e-13)
    assert ok, f"EHVI at explicit points: {worst:.3g} x tolerance"


@pytest.mark.parametrize("name,n_obj", [("welded_like", 2), ("mixed_cont", 2), ("mixed_cont", 3)])
def test_mc_pr_ehvi_vs_oracle(name, n_obj):
    c, g0, oacq, acq, pr = _build_ehvi(name, n_obj)
    space = oracle_space(g0)
    theta = g0["theta"]
    go = torch.linspace(0.6, 1.4, c["b"], dtype=torch.float64)
    pre = (2.0, 0.01) if c["grad_estimator"] == "reinforce_ma" else (0.0, 0.0)
    ref = P.mc_pr(space, theta, oacq, g0["base_samples"], g0["base_samples_categorical"],
                  grad_estimator=c["grad_estimator"], ma_counter=pre[0], ma_hidden=pre[1], grad_output=go)
    pr._ma_state.copy_(torch.tensor(pre, dtype=torch.float64))
    X = theta.to(DEV).requires_grad_(True)
    val = pr(X)
    grad = torch.autograd.grad(val, X, grad_outputs=go.to(DEV))[0]
    ok, w = close(val, ref.value, 1e-9, 1e-13)
    assert ok, f"PR-EHVI value: {w:.3g} x tolerance"
    ok, w = close(grad, ref.grad, 1e-9, 1e-12)
    assert ok, f"PR-EHVI gradient (score function + pathwise through every objective): {w:.3g} x tolerance"
    lo = torch.zeros(c["dim"], dtype=torch.float64, device=DEV)
    hi = torch.ones(c["dim"], dtype=torch.float64, device=DEV)
    with torch.no_grad():
        v0 = pr(theta.to(DEV))
    cand, v1 = B.gen_candidates_torch(theta.to(DEV), pr, lo, hi, options={"maxiter": 10})
    assert float(v1.sum()) >= float(v0.sum()) - 1e-9
    chosen, value = B.finalize_candidates(pr, cand, seed=2)
    assert chosen.shape == (1, c["dim"]) and float(value) >= 0.0
