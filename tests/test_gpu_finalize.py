"""Candidate finalisation on the device:
  * sample_candidates against the draws the REFERENCE made under a recorded torch seed (golden fixtures),
  * the Philox / explicit-uniform kernel against the oracle's restatement, bit for bit,
  * prb_finalize_candidates (sample -> base-AF re-scoring -> argmax, one enqueue) against the oracle,
  * the trust-region box kernel against the host restatement of run_one_replication.py:410-440,
  * optimize_acqf_mixed on the device against a brute-force enumeration."""
import ctypes as C

import pytest
import torch

import bo_pr_b200 as B
from bo_pr_b200 import _capi as capi
from oracle import pr_oracle as P
from oracle.philox import philox_uniform
from tests.golden_util import ANALYTIC_CASES, MC_CASES, load, oracle_acq, oracle_space
from tests.product_util import close, product_acq, product_analytic_pr, product_mc_pr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pr(g, name):
    return product_analytic_pr(g, DEV) if name in ANALYTIC_CASES else product_mc_pr(g, DEV)


@pytest.mark.parametrize("name", MC_CASES + ANALYTIC_CASES)
def test_sample_candidates_reproduces_reference_draws(name):
    """rng="torch": probabilities from the CUDA sampler, draws from torch's generator in the reference's order.  With the
    draws made on the CPU generator the rows equal what the reference itself produced under the same seed (the fixtures
    were recorded from a CPU run of /root/reference)."""
    g = load(name)
    rec = g["sample_candidates"]
    pr = _pr(g, name)
    torch.manual_seed(rec["seed"])
    out = pr.sample_candidates(rec["X"].to(DEV), rng="torch", draw_device="cpu")
    assert out.device.type == "cuda"
    assert torch.equal(out.cpu(), rec["out"]), int((out.cpu() != rec["out"]).sum())
    # default draw device (the CUDA generator): a valid configuration, continuous columns untouched, repeatable per seed
    torch.manual_seed(5)
    a = pr.sample_candidates(rec["X"].to(DEV))
    torch.manual_seed(5)
    b = pr.sample_candidates(rec["X"].to(DEV))
    assert torch.equal(a, b)
    cont = pr.cont_indices.cpu()
    assert torch.equal(a.cpu()[..., cont], rec["X"][..., cont])


@pytest.mark.parametrize("name", MC_CASES + ANALYTIC_CASES)
def test_sample_candidates_kernel_vs_oracle(name):
    """rng="philox" (one kernel, counter = (restart, parameter)) and explicit uniforms: bit-exact against the oracle fed
    the same uniforms."""
    g = load(name)
    space = oracle_space(g)
    X = g["sample_candidates"]["X"]
    b = X.shape[0]
    nu = space.n_int + space.n_cat
    pr = _pr(g, name)
    seed, offset = 12345, (1 << 32) + 7
    u = torch.from_numpy(philox_uniform(seed, offset, b, nu))
    ref = P.sample_candidates(space, X, u=u)
    out = pr.sample_candidates(X.to(DEV), rng="philox", seed=seed, offset=offset)
    assert torch.equal(out.cpu(), ref), int((out.cpu() != ref).sum())
    gen = torch.Generator().manual_seed(9)
    u2 = torch.rand(b, nu, generator=gen, dtype=torch.float64)
    u2[0] = 0.0
    u2[-1] = 1.0 - 2.0 ** -53  # the inverse-CDF fallback (no cumsum entry exceeds u)
    out2 = pr.sample_candidates(X.to(DEV), u=u2.to(DEV))
    assert torch.equal(out2.cpu(), P.sample_candidates(space, X, u=u2))


@pytest.mark.parametrize("name", MC_CASES + ANALYTIC_CASES)
def test_finalize_candidates_vs_oracle(name):
    """prb_finalize_candidates: sampled rows == the sampler kernel's, values == the oracle's base AF on those rows
    (numeric layout where the model wants it), best row == argmax."""
    g = load(name)
    c = g["case"]
    space, oacq = oracle_space(g), oracle_acq(g)
    X = g["sample_candidates"]["X"]
    pr = _pr(g, name)
    chosen, value, rows, vals = B.finalize_candidates(pr, X.to(DEV), seed=3, return_all=True)
    ref_rows = pr.sample_candidates(X.to(DEV), rng="philox", seed=3)
    assert torch.equal(rows, ref_rows)
    xo = rows.cpu()
    xo = P.one_hot_to_numeric(space, xo) if c["apply_numeric"] else xo
    with torch.no_grad():
        ref_vals = oacq(xo)
    ok, worst = close(vals, ref_vals, 1e-9, 1e-12)
    assert ok, f"re-scored values: {worst:.3g} x tolerance"
    i = int(torch.argmax(vals))
    assert torch.equal(chosen, rows[i]) and float(value) == float(vals[i])


def test_trust_region_bounds_kernel_vs_host():
    from bo_pr_b200.trust_region import TurboState, trust_region_bounds
    gen = torch.Generator().manual_seed(0)
    d, n = 16, 57
    X = torch.rand(n, d, generator=gen, dtype=torch.float64)
    std = torch.stack([torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)])
    for constrained in (False, True):
        for feas_shift in (0.0, -10.0):  # with -10 no point is feasible: least-violating centre
            Y = torch.randn(n, 3 if constrained else 1, generator=gen, dtype=torch.float64)
            if constrained:
                Y[:, 1:] += feas_shift
            st = TurboState(dim=d, batch_size=1, is_constrained=constrained, length=0.4)
            ref = trust_region_bounds(st, X, Y, std, binary_dims=[8, 9], categorical_start=14)
            out = trust_region_bounds(st, X.to(DEV), Y.to(DEV), std.to(DEV), binary_dims=[8, 9], categorical_start=14)
            assert out.device.type == "cuda"
            assert torch.equal(out.cpu(), ref)


def test_optimize_acqf_mixed_on_device_vs_enumeration():
    """a14: optimize_acqf_mixed over the four settings of two binary columns of a base EI on the GPU: the winner carries
    the pinned columns, its value is the device AF at that point (checked against the oracle), and it is at least as good
    as every initial condition under every setting."""
    g = load("mixed_int_f1_ei")
    acq = product_acq(g, DEV)
    oacq = oracle_acq(g)
    d = g["case"]["dim"]
    bounds = torch.stack([torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)]).to(DEV)
    ff_list = [{8: float(a), 9: float(b)} for a in (0, 1) for b in (0, 1)]
    gen = torch.Generator().manual_seed(1)
    ics = torch.rand(6, 1, d - 2, generator=gen, dtype=torch.float64).to(DEV)  # free columns only (the reduced space)
    cand, val = B.optimize_acqf_mixed(acq, bounds, q=1, num_restarts=6, fixed_features_list=ff_list,
                                      batch_initial_conditions=ics, options={"maxiter": 40})
    assert cand.shape == (1, d) and cand.device.type == "cuda"
    assert (float(cand[0, 8]), float(cand[0, 9])) in [(a, b) for a in (0.0, 1.0) for b in (0.0, 1.0)]
    with torch.no_grad():
        ref_val = oacq(cand.cpu().unsqueeze(0))
    ok, worst = close(val.reshape(-1), ref_val.reshape(-1), 1e-9, 1e-12)
    assert ok, f"value at the returned candidate: {worst:.3g} x tolerance"
    # L-BFGS-B never returns a point worse than where it started: the winner beats every initial condition in every setting
    free = [i for i in range(d) if i not in (8, 9)]
    best_start = -float("inf")
    for ff in ff_list:
        Xf = torch.empty(6, 1, d, dtype=torch.float64, device=DEV)
        Xf[..., free] = ics
        for col, v in ff.items():
            Xf[..., col] = v
        with torch.no_grad():
            best_start = max(best_start, float(acq(Xf).max()))
    assert float(val) >= best_start - 1e-12
