"""GPU: the fused on-device multi-start Adam (prb_mc_adam: gen.py:304-343 semantics in one call) against the host-driven
loop over the same CUDA value+gradient path and against the CPU oracle's restatement; optimize_acqf end to end."""
import pytest
import torch

from bench_workload import make_problem, make_theta, sobol_base_samples
from oracle import pr_oracle as P
from tests.golden_util import load, oracle_acq, oracle_space
from tests.product_util import product_mc_pr

import bo_pr_b200 as B

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bounds(d):
    return torch.stack([torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)]).to(DEV)


@pytest.mark.parametrize("name", ["ackley13_ei", "rosenbrock10_ucb", "mixed_int_f1_ei", "chemistry_mc_ei",
                                  "chemistry_latent_mc_ei"])
def test_fused_adam_matches_host_loop_and_oracle(name):
    g = load(name)
    c = g["case"]
    d = c["dim"]
    ics = g["theta"].to(DEV)
    bounds = _bounds(d)
    steps = 12
    # (1) fused: one prb_mc_adam call
    pr_f = product_mc_pr(g, DEV)
    cand_f, val_f = B.gen_candidates_torch(ics, pr_f, bounds[0], bounds[1], options={"maxiter": steps})
    # (2) host-driven loop over the same class (torch.optim.Adam, one C-ABI call per step)
    pr_h = product_mc_pr(g, DEV)
    cand_h, val_h = B.gen_candidates_torch(ics, pr_h, bounds[0], bounds[1], options={"maxiter": steps, "fused": False})
    assert torch.allclose(cand_f, cand_h, rtol=0, atol=1e-10), (cand_f - cand_h).abs().max()
    assert torch.allclose(val_f, val_h, rtol=1e-8, atol=1e-12)
    assert torch.allclose(pr_f._ma_state, pr_h._ma_state, rtol=1e-9, atol=0)
    assert float(pr_f._ma_state[0]) == steps + 1  # every forward (steps + final evaluation) advanced the EMA counter
    # (3) CPU oracle: reference semantics end to end
    space, oacq = oracle_space(g), oracle_acq(g)
    st = {"c": 0.0, "h": 0.0}

    def vg(X):
        r = P.mc_pr(space, X, oacq, g["base_samples"], g["base_samples_categorical"],
                    grad_estimator=c["grad_estimator"], ma_counter=st["c"], ma_hidden=st["h"],
                    apply_numeric=c["apply_numeric"])
        st["c"], st["h"] = r.ma_counter, r.ma_hidden
        return r.value, r.grad

    def vo(X):
        r = P.mc_pr(space, X, oacq, g["base_samples"], g["base_samples_categorical"],
                    grad_estimator=c["grad_estimator"], ma_counter=st["c"], ma_hidden=st["h"], with_grad=False,
                    apply_numeric=c["apply_numeric"])
        st["c"], st["h"] = r.ma_counter, r.ma_hidden
        return r.value

    theta_o, val_o = P.adam_multistart(vg, vo, g["theta"], torch.zeros(d, dtype=torch.float64),
                                       torch.ones(d, dtype=torch.float64), maxiter=steps)
    assert torch.allclose(cand_f.cpu(), theta_o, rtol=0, atol=1e-8), (cand_f.cpu() - theta_o).abs().max()
    assert torch.allclose(val_f.cpu(), val_o, rtol=1e-7, atol=1e-11)


def test_optimize_acqf_pr_ei_improves_over_initial_conditions():
    """The reference's call (run_one_replication.py:494-501) at the ackley13 shape: restarts in batches of 5."""
    g = make_problem(119, seed=3)
    g["case"].update(mc=128, grad_estimator="reinforce_ma")
    g["base_samples"] = sobol_base_samples(128)
    g["base_samples_categorical"] = None
    pr = product_mc_pr(g, DEV)
    bounds = _bounds(13)
    torch.manual_seed(0)
    opts = {"batch_limit": 5, "init_batch_limit": 32, "maxiter": 200, "nonnegative": True, "seed": 7}
    ics = B.gen_batch_initial_conditions(pr, bounds, q=1, num_restarts=20, raw_samples=1024, options=opts)
    assert float(pr._ma_state[0]) == 32  # 1024 / 32 screening calls advanced the EMA counter
    with torch.no_grad():
        v0 = pr(ics)
    cand, vals = B.optimize_acqf(pr, bounds, q=1, num_restarts=20, options=opts, batch_initial_conditions=ics,
                                 stochastic=True, return_best_only=False)
    assert cand.shape == (20, 1, 13) and vals.shape == (20,)
    assert bool((cand >= 0).all()) and bool((cand <= 1).all())
    assert float(vals.max()) >= float(v0.max()) - 1e-12
    assert float(vals.mean()) > float(v0.mean())
    # candidates finalisation (probabilistic_reparameterization.py:198-233): valid discrete points
    final = pr.sample_candidates(cand)
    assert bool(((final[..., 3:] == 0) | (final[..., 3:] == 1)).all())
    assert torch.equal(final[..., :3], cand[..., :3])


def test_screen_equals_chunked_forward_calls():
    """prb_mc_screen (the init_batch_limit loop inside the library) = the chunked Python loop, EMA state included."""
    g = make_problem(100, seed=4)
    g["case"].update(mc=64, grad_estimator="reinforce_ma")
    g["base_samples"] = sobol_base_samples(64)
    g["base_samples_categorical"] = None
    X = make_theta(70, seed=8).to(DEV)  # ragged last chunk
    pr_a = product_mc_pr(g, DEV)
    with torch.no_grad():
        ya = torch.cat([pr_a(X[s:s + 32]) for s in range(0, 70, 32)])
    pr_b = product_mc_pr(g, DEV)
    yb = pr_b.screen(X, 32)
    assert torch.allclose(ya, yb, rtol=1e-12, atol=1e-15)
    assert torch.allclose(pr_a._ma_state, pr_b._ma_state, rtol=1e-12, atol=0)
    assert float(pr_b._ma_state[0]) == 3.0


@pytest.mark.parametrize("b,mc,n", [(40, 128, 100), (3, 32, 127), (6, 96, 60)])
def test_small_step_paths_match_host_loop(b, mc, n):
    """Experiment-sized problems: the persistent cooperative kernel (restarts x mc / 32 <= 148 CTAs), the one-kernel step
    replayed from a graph (more CTAs than SMs) and the host-driven loop over single fused calls follow one trajectory."""
    g = make_problem(n, seed=6)
    g["case"].update(mc=mc, grad_estimator="reinforce_ma")
    g["base_samples"] = sobol_base_samples(mc)
    g["base_samples_categorical"] = None
    bounds = _bounds(13)
    ics = make_theta(b, seed=5).to(DEV)
    pr_f = product_mc_pr(g, DEV)
    cand_f, val_f = B.gen_candidates_torch(ics, pr_f, bounds[0], bounds[1], options={"maxiter": 8})
    pr_h = product_mc_pr(g, DEV)
    cand_h, val_h = B.gen_candidates_torch(ics, pr_h, bounds[0], bounds[1], options={"maxiter": 8, "fused": False})
    assert torch.allclose(cand_f, cand_h, rtol=0, atol=1e-10), (cand_f - cand_h).abs().max()
    assert torch.allclose(val_f, val_h, rtol=1e-8, atol=1e-12)
    assert torch.allclose(pr_f._ma_state, pr_h._ma_state, rtol=1e-9, atol=0)


@pytest.mark.parametrize("name,n,mc", [("rosenbrock10_ucb", 150, 64), ("mixed_int_f1_ei", 131, 48), ("rosenbrock10_ucb", 90, 40)])
def test_generic_fused_step_matches_host_loop_and_oracle(name, n, mc):
    """Kernels that do not separate, outside the single-kernel limits (n + 1 > 128 or mc not a multiple of 32): the
    six-launch fused step replayed from a graph, the host-driven loop and the oracle follow one trajectory."""
    from tests.random_cases import build_problem

    c = dict(load(name)["case"], n_train=n, mc=mc, b=4, seed=11)
    g = build_problem(c)
    d = c["dim"]
    bounds = _bounds(d)
    ics = g["theta"].to(DEV)
    steps = 6
    pr_f = product_mc_pr(g, DEV)
    cand_f, val_f = B.gen_candidates_torch(ics, pr_f, bounds[0], bounds[1], options={"maxiter": steps})
    pr_h = product_mc_pr(g, DEV)
    cand_h, val_h = B.gen_candidates_torch(ics, pr_h, bounds[0], bounds[1], options={"maxiter": steps, "fused": False})
    assert torch.allclose(cand_f, cand_h, rtol=0, atol=1e-10), (cand_f - cand_h).abs().max()
    assert torch.allclose(val_f, val_h, rtol=1e-8, atol=1e-12)
    assert torch.allclose(pr_f._ma_state, pr_h._ma_state, rtol=1e-9, atol=0)
    space, oacq = oracle_space(g), oracle_acq(g)
    st = {"c": 0.0, "h": 0.0}

    def vg(X, with_grad=True):
        r = P.mc_pr(space, X, oacq, g["base_samples"], g["base_samples_categorical"],
                    grad_estimator=c["grad_estimator"], ma_counter=st["c"], ma_hidden=st["h"], with_grad=with_grad)
        st["c"], st["h"] = r.ma_counter, r.ma_hidden
        return (r.value, r.grad) if with_grad else r.value

    theta_o, val_o = P.adam_multistart(vg, lambda X: vg(X, False), g["theta"], torch.zeros(d, dtype=torch.float64),
                                       torch.ones(d, dtype=torch.float64), maxiter=steps)
    assert torch.allclose(cand_f.cpu(), theta_o, rtol=0, atol=1e-8), (cand_f.cpu() - theta_o).abs().max()
    assert torch.allclose(val_f.cpu(), val_o, rtol=1e-7, atol=1e-11)


@pytest.mark.parametrize("name,n", [("chemistry_analytic_ei", None), ("int_cat_analytic_pm", None), ("chemistry_analytic_ei", 140)])
def test_fused_analytic_adam_matches_host_loop_and_oracle(name, n):
    """Analytic PR (enumeration): prb_analytic_adam -- the seven-launch fused step replayed from a CUDA graph (six launches
    with a single row tile) -- against the host-driven torch.optim.Adam loop over the same class and against the oracle."""
    from tests.product_util import product_analytic_pr
    from tests.random_cases import build_problem

    g = load(name)
    if n is not None:  # a second row tile (n + 1 > 128): separate acquisition epilogue
        g = build_problem(dict(g["case"], n_train=n, b=3, mc=8, seed=5))
    c = g["case"]
    d = c["dim"]
    bounds = _bounds(d)
    ics = g["theta"].to(DEV)
    steps = 7
    pr_f = product_analytic_pr(g, DEV)
    cand_f, val_f = B.gen_candidates_torch(ics, pr_f, bounds[0], bounds[1], options={"maxiter": steps})
    pr_h = product_analytic_pr(g, DEV)
    cand_h, val_h = B.gen_candidates_torch(ics, pr_h, bounds[0], bounds[1], options={"maxiter": steps, "fused": False})
    assert torch.allclose(cand_f, cand_h, rtol=0, atol=1e-10), (cand_f - cand_h).abs().max()
    assert torch.allclose(val_f, val_h, rtol=1e-8, atol=1e-12)
    space, oacq = oracle_space(g), oracle_acq(g)

    def vg(X):
        v, gr, _, _ = P.analytic_pr(space, X, oacq, apply_numeric=c["apply_numeric"])
        return v, gr

    def vo(X):
        return P.analytic_pr(space, X, oacq, apply_numeric=c["apply_numeric"], with_grad=False)[0]

    theta_o, val_o = P.adam_multistart(vg, vo, g["theta"], torch.zeros(d, dtype=torch.float64),
                                       torch.ones(d, dtype=torch.float64), maxiter=steps)
    assert torch.allclose(cand_f.cpu(), theta_o, rtol=0, atol=1e-8), (cand_f.cpu() - theta_o).abs().max()
    assert torch.allclose(val_f.cpu(), val_o, rtol=1e-7, atol=1e-11)
