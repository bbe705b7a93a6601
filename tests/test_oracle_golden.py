"""The oracle's L1 restatement (oracle/pr_oracle.py) against golden vectors produced by the REFERENCE's own code
(oracle/gen_golden.py imported /root/reference unmodified).  CPU only."""
import pytest
import torch

from oracle import pr_oracle as P
from tests.golden_util import ANALYTIC_CASES, MC_CASES, load, oracle_acq, oracle_space


@pytest.mark.parametrize("name", MC_CASES)
def test_mc_pr_matches_reference(name):
    g = load(name)
    c = g["case"]
    space = oracle_space(g)
    acq = oracle_acq(g)
    for rec in g["calls"]:
        res = P.mc_pr(space, rec["X"], acq, g["base_samples"], g["base_samples_categorical"],
                      grad_estimator=c["grad_estimator"], ma_counter=rec["ma_pre"][0], ma_hidden=rec["ma_pre"][1],
                      apply_numeric=c["apply_numeric"], grad_output=rec["grad_output"])
        # discrete samples: bit-exact
        assert torch.equal(res.tilde_x, rec["tilde_x"])
        assert torch.equal(res.prob, rec["prob"])
        # values / gradients: same arithmetic, same library -> essentially exact
        torch.testing.assert_close(res.acq_values, rec["acq_values"], rtol=1e-13, atol=1e-15)
        torch.testing.assert_close(res.value, rec["value"], rtol=1e-13, atol=1e-15)
        torch.testing.assert_close(res.grad, rec["grad"], rtol=1e-12, atol=1e-14)
        assert res.ma_counter == rec["ma_post"][0]
        assert abs(res.ma_hidden - rec["ma_post"][1]) <= 1e-15 * max(1.0, abs(rec["ma_post"][1]))


@pytest.mark.parametrize("name", ANALYTIC_CASES)
def test_analytic_pr_matches_reference(name):
    g = load(name)
    c = g["case"]
    space = oracle_space(g)
    acq = oracle_acq(g)
    rec = g["calls"][0]
    table = P.all_discrete_options(space)
    assert torch.equal(table, g["all_discrete_options"].to(table.dtype))
    value, grad, acq_values, probs = P.analytic_pr(space, g["theta"], acq, apply_numeric=c["apply_numeric"])
    torch.testing.assert_close(probs, rec["probs"], rtol=1e-14, atol=1e-16)
    torch.testing.assert_close(value, rec["value"], rtol=1e-13, atol=1e-15)
    torch.testing.assert_close(grad, rec["grad"], rtol=1e-11, atol=1e-13)


def test_first_backward_equals_plain_reinforce():
    """with ma_counter == 0 the reinforce_ma baseline is 0."""
    g = load("ackley13_ei")
    space, acq = oracle_space(g), oracle_acq(g)
    rec = g["calls"][0]
    a = P.mc_pr(space, rec["X"], acq, g["base_samples"], None, grad_estimator="reinforce_ma")
    b = P.mc_pr(space, rec["X"], acq, g["base_samples"], None, grad_estimator="reinforce")
    assert torch.equal(a.grad, b.grad)


def test_analytic_probs_sum_to_one_inside_bounds():
    g = load("chemistry_analytic_ei")
    space = oracle_space(g)
    probs = P.analytic_probs(space, g["theta"], P.all_discrete_options(space))
    torch.testing.assert_close(probs.sum(-1), torch.ones(probs.shape[0], dtype=torch.float64), rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", MC_CASES + ANALYTIC_CASES)
def test_sample_candidates_matches_reference_draws(name):
    """a12: the final discrete draw per restart under the recorded torch seed -- same generator consumption order as the
    reference (one Bernoulli call per integer parameter, one Categorical call per one-hot block) => identical rows."""
    g = load(name)
    rec = g["sample_candidates"]
    torch.manual_seed(rec["seed"])
    out = P.sample_candidates(oracle_space(g), rec["X"])
    assert torch.equal(out, rec["out"])
