"""GPU parity: the CUDA path (through the C ABI, called from the reference-API classes) against
  (1) the golden vectors the REFERENCE's own code produced (tests/golden/*.pt, oracle/gen_golden.py), and
  (2) the CPU oracle on seeded inputs at sizes it finishes in seconds.
Tolerance (BASELINE.json north_star): sampled discrete configurations bit-exact; values / gradients 1e-9 relative
(float64) with an absolute floor of 1e-12 (sigma^2 and EI cancel near training points)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import gp_oracle as G
from oracle import pr_oracle as P
from oracle.philox import philox_uniform
from tests.golden_util import (ANALYTIC_CASES, MC_CASES, load, oracle_acq, oracle_latent_transform, oracle_space,
                               oracle_spec)
from tests.product_util import (close, integer_bounds, product_acq, product_analytic_pr, product_mc_pr,
                                product_model)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _assert_close(a, b, what, rtol=1e-9, atol=1e-12):
    ok, worst = close(a, b, rtol, atol)
    assert ok, f"{what}: worst error = {worst:.3g} x tolerance"


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", MC_CASES + ANALYTIC_CASES)
def test_kernel_matrix_and_posterior(name):
    """a9/a10: K(x*, X) and the posterior (mean, variance) against the L0 oracle."""
    g = load(name)
    model = product_model(g, DEV)
    tf = oracle_latent_transform(g)  # mixed_latent: the model's inputs are the numeric layout, embedded by the model
    ogp = G.OracleGP(g["train_X_model"], g["train_Y"], oracle_spec(g), g["noise"], g["mean_const"], input_transform=tf)
    gen = torch.Generator().manual_seed(5)
    Xk = g["train_X_model"] if tf is None else g["train_X_numeric"]
    # test points: perturbed training points (hard: small variance) and fresh points with valid discrete codes
    idx = torch.randint(0, Xk.shape[0], (64,), generator=gen)
    Xt = Xk[idx].clone()
    cont = g["cont_dims"]
    Xt[:, cont] = (Xt[:, cont] + 0.2 * torch.randn(64, len(cont), generator=gen)).clamp(0, 1)
    Xt[:8, cont] = Xk[idx[:8]][:, cont] + 1e-3
    post = model.posterior(Xt.to(DEV).unsqueeze(-2))
    mean_o, var_o = ogp.posterior(Xt)
    _assert_close(post.mean.reshape(-1), mean_o, "posterior mean")
    _assert_close(post.variance.reshape(-1), var_o, "posterior variance")
    # the state's kernel matrix
    st = model.state
    lib = st.lib
    K = torch.empty(64, st.n, dtype=torch.float64, device=DEV)
    Xt_k = Xt if tf is None else tf(Xt)
    xt = Xt_k.to(DEV).contiguous()
    rc = lib.prb_kernel_matrix(C.byref(st._desc), xt.data_ptr(), 64, K.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    _assert_close(K, G.eval_kernel(oracle_spec(g), Xt_k, g["train_X_model"]), "K(x*, X)", rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("name", MC_CASES)
def test_base_acq_value_and_gradient(name):
    """a11 + pathwise gradient: acq(X) and d acq / dX against oracle autograd."""
    g = load(name)
    acq = product_acq(g, DEV)
    oacq = oracle_acq(g)
    gen = torch.Generator().manual_seed(6)
    Xk = g["train_X_model"] if g.get("latent") is None else g["train_X_numeric"]
    idx = torch.randint(0, Xk.shape[0], (48,), generator=gen)
    Xt = Xk[idx].clone()
    cont = g["cont_dims"]
    Xt[:, cont] = (Xt[:, cont] + 0.15 * torch.randn(48, len(cont), generator=gen)).clamp(0, 1)
    Xo = Xt.clone().unsqueeze(-2).requires_grad_(True)
    vo = oacq(Xo)
    go = torch.autograd.grad(vo.sum(), Xo)[0]
    Xp = Xt.to(DEV).unsqueeze(-2).requires_grad_(True)
    vp = acq(Xp)
    gp = torch.autograd.grad(vp.sum(), Xp)[0]
    _assert_close(vp, vo, "acq value")
    # only columns that feed a Matern factor carry a gradient on either side (categorical codes: 0 on both)
    _assert_close(gp[..., cont], go[..., cont], "acq gradient (continuous columns)", atol=1e-11)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", MC_CASES)
def test_sampler_bit_exact_vs_reference(name):
    """a2/a3/a5: sampled configurations (bit-exact), rounding components, probabilities -- vs the reference's outputs."""
    g = load(name)
    c = g["case"]
    pr = product_mc_pr(g, DEV)
    lib = pr.input_transform["round"].lib
    sp = pr._space()
    rnd = pr.input_transform["round"]
    ui, uc = rnd.base_sample_ptrs()
    mismatches = 0
    for rec in g["calls"]:
        X = rec["X"].to(DEV)
        b, d = X.shape[0], X.shape[-1]
        mc = c["mc"]
        n_disc = rec["prob"].shape[-1]
        tilde = torch.empty(b, mc, d, dtype=torch.float64, device=DEV)
        z = torch.empty(b, mc, n_disc, dtype=torch.uint8, device=DEV)
        prob = torch.empty(b, n_disc, dtype=torch.float64, device=DEV)
        th = X.reshape(b, d).contiguous()
        rc = lib.prb_pr_sample(C.byref(sp), th.data_ptr(), b, mc, None if ui is None else ui.data_ptr(),
                               None if uc is None else uc.data_ptr(), 0, 0, tilde.data_ptr(), z.data_ptr(),
                               prob.data_ptr(), None)
        assert rc == 0
        torch.cuda.synchronize()
        ref_tilde = rec["tilde_x"].reshape(b, mc, d)
        mismatches += int((tilde.cpu() != ref_tilde).sum())
        # probabilities: exp() differs in the last ulp between SLEEF (torch CPU) and CUDA libdevice
        _assert_close(prob, rec["prob"].reshape(b, n_disc), "rounding probabilities", rtol=1e-14, atol=1e-300)
        # z as the reference's Function derives it (probabilistic_reparameterization.py:454-465)
        space = oracle_space(g)
        res = P.mc_pr(space, rec["X"], lambda t: t.sum(dim=(-1, -2)), g["base_samples"], g["base_samples_categorical"],
                      with_grad=False)
        assert torch.equal(z.cpu().to(torch.float64), res.z.reshape(b, mc, n_disc))
        # through the reference-API transform chain too (un-normalised space in, normalised out)
        out = pr.input_transform(X.unsqueeze(-3))
        assert torch.equal(out.cpu(), rec["tilde_x"])
    assert mismatches == 0, f"{mismatches} sampled entries differ from the reference"


@pytest.mark.parametrize("name", MC_CASES)
def test_mc_pr_value_grad_vs_reference(name):
    """a6/a7: MC-PR value, gradient and EMA state, call by call, against the reference's recorded outputs."""
    g = load(name)
    pr = product_mc_pr(g, DEV)
    for rec in g["calls"]:
        pre = torch.tensor(rec["ma_pre"], dtype=torch.float64, device=DEV)
        assert torch.allclose(pr._ma_state, pre, rtol=1e-12, atol=0), "EMA state drifted before the call"
        pr._ma_state.copy_(pre)  # replay from the reference's exact pre-call state
        X = rec["X"].to(DEV).requires_grad_(True)
        val = pr(X)
        grad = torch.autograd.grad(val, X, grad_outputs=rec["grad_output"].to(DEV))[0]
        _assert_close(val, rec["value"], "PR value")
        _assert_close(grad, rec["grad"], "PR gradient", atol=1e-11)
        post = torch.tensor(rec["ma_post"], dtype=torch.float64)
        _assert_close(pr._ma_state, post, "EMA state after the call", rtol=1e-12)


@pytest.mark.parametrize("name", MC_CASES)
def test_mc_pr_per_sample_values(name):
    g = load(name)
    c = g["case"]
    pr = product_mc_pr(g, DEV)
    st = pr.acq_function.state
    rec = g["calls"][0]
    X = rec["X"].to(DEV)
    b, d, mc = X.shape[0], X.shape[-1], c["mc"]
    rnd = pr.input_transform["round"]
    ui, uc = rnd.base_sample_ptrs()
    from bo_pr_b200.acquisition import acq_desc_of
    sp, aq = pr._space(), acq_desc_of(pr.acq_function)
    value = torch.empty(b, dtype=torch.float64, device=DEV)
    av = torch.empty(b, mc, dtype=torch.float64, device=DEV)
    ws = st.workspace(b * mc, d, b)
    th = X.reshape(b, d).contiguous()
    rc = st.lib.prb_mc_value_grad(st.handle, C.byref(sp), C.byref(aq), th.data_ptr(), b, mc,
                                  None if ui is None else ui.data_ptr(), None if uc is None else uc.data_ptr(), 0, 0,
                                  0, None, 0, None, value.data_ptr(), None, av.data_ptr(), ws.data_ptr(), ws.numel(), 0,
                                  None)
    assert rc == 0
    torch.cuda.synchronize()
    _assert_close(av, rec["acq_values"], "per-sample acquisition values")
    _assert_close(value, rec["value"], "PR value (value-only entry)")


@pytest.mark.parametrize("name", ANALYTIC_CASES)
def test_analytic_pr_vs_reference(name):
    """a4/a8: enumeration table, P(z | theta), analytic PR value and gradient against the reference's outputs."""
    g = load(name)
    pr = product_analytic_pr(g, DEV)
    rnd = pr.input_transform["round"]
    assert torch.equal(rnd.all_discrete_options.cpu(), g["all_discrete_options"].to(torch.float64))
    rec = g["calls"][0]
    theta = g["theta"].to(DEV)
    xun = pr.input_transform["unnormalize"](theta) if "unnormalize" in pr.input_transform else theta
    probs = rnd.get_probs(xun).squeeze(-1)
    _assert_close(probs, rec["probs"], "P(z | theta)", rtol=1e-13, atol=1e-300)
    X = theta.clone().requires_grad_(True)
    val = pr(X)
    grad = torch.autograd.grad(val.sum(), X)[0]
    _assert_close(val, rec["value"], "analytic PR value")
    _assert_close(grad, rec["grad"], "analytic PR gradient", atol=1e-11)
    # the pasted inputs (transform): discrete columns from the table, continuous columns from theta
    out = rnd(xun.unsqueeze(-3))
    table = g["all_discrete_options"].to(torch.float64)
    space = oracle_space(g)
    assert torch.equal(out[0, :, 0, space.discrete_indices].cpu(), table[:, space.discrete_indices])
    assert torch.equal(out[:, 0, 0, space.cont_indices].cpu(), g["theta"][:, 0, space.cont_indices])


# ---------------------------------------------------------------------------------------------------------------------
def test_philox_stream_bit_exact():
    from bo_pr_b200 import _capi as capi
    lib = capi.load_library()
    for seed, off, mc, nc in [(0, 0, 128, 10), (2**40 + 17, 2**33 + 5, 257, 13), (123456789, 1, 1024, 3)]:
        out = torch.empty(mc, nc, dtype=torch.float64, device=DEV)
        assert lib.prb_philox_uniform(seed, off, mc, nc, out.data_ptr(), None) == 0
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), philox_uniform(seed, off, mc, nc))


def test_philox_fused_sampler_equals_materialised_stream():
    """The sampler's on-the-fly Philox draws (u_int = u_cat = NULL) give the same configurations as feeding it the
    materialised stream -- and as the oracle consuming the numpy restatement of that stream."""
    g = load("neg_int_onehot_ei")
    c = g["case"]
    pr = product_mc_pr(g, DEV)
    lib = pr.input_transform["round"].lib
    sp = pr._space()
    X = g["calls"][0]["X"].to(DEV)
    b, d, mc = X.shape[0], X.shape[-1], 96
    seed, off = 99, 4
    n_int, n_cat = len(c["integer_indices"]), len(c["categorical_features"])
    u = torch.from_numpy(philox_uniform(seed, off, mc, n_int + n_cat))
    ui = u[:, :n_int].contiguous().to(DEV)
    uc = u[:, n_int:].contiguous().to(DEV)
    th = X.reshape(b, d).contiguous()
    t1 = torch.empty(b, mc, d, dtype=torch.float64, device=DEV)
    t2 = torch.empty_like(t1)
    assert lib.prb_pr_sample(C.byref(sp), th.data_ptr(), b, mc, None, None, seed, off, t1.data_ptr(), None, None, None) == 0
    assert lib.prb_pr_sample(C.byref(sp), th.data_ptr(), b, mc, ui.data_ptr(), uc.data_ptr(), 0, 0, t2.data_ptr(), None,
                             None, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(t1, t2)
    space = oracle_space(g)
    tilde = P.normalize(space, P.mc_round(space, P.unnormalize(space, g["calls"][0]["X"].unsqueeze(-3)),
                                          u[:, :n_int].reshape(mc, 1, n_int), u[:, n_int:].reshape(mc, 1, n_cat)))
    assert torch.equal(t1.cpu(), tilde.reshape(b, mc, d))
