"""The synthetic sweep shape (BASELINE.json configs[4]): CUDA path vs the CPU oracle at sizes the oracle finishes in
seconds, and size-independent properties at the full bench size (n = 1000, 64 restarts x 1024 MC samples)."""
import ctypes as C

import pytest
import torch

from bench_workload import make_problem, make_theta, sobol_base_samples
from oracle import pr_oracle as P
from tests.golden_util import oracle_acq, oracle_space
from tests.product_util import close, product_acq, product_mc_pr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _problem(n, mc, acq="ei", est="reinforce_ma", seed=0):
    g = make_problem(n, seed=seed)
    g["case"].update(acq=acq, mc=mc, grad_estimator=est)
    g["base_samples"] = sobol_base_samples(mc)
    g["base_samples_categorical"] = None
    return g


def _assert_close(a, b, what, rtol=1e-9, atol=1e-12):
    ok, worst = close(a, b, rtol, atol)
    assert ok, f"{what}: worst error = {worst:.3g} x tolerance"


@pytest.mark.parametrize("n,b,mc,acq", [(200, 8, 128, "ei"), (333, 5, 96, "ucb"), (1000, 4, 256, "ei"),
                                        (1537, 3, 64, "ei"),
                                        (2000, 8, 128, "ei")])  # 16 row tiles x 1 024 columns: the 32-column tile choice
def test_value_grad_vs_oracle(n, b, mc, acq):
    g = _problem(n, mc, acq)
    space, oacq = oracle_space(g), oracle_acq(g)
    theta = make_theta(b)
    go = torch.linspace(0.5, 1.5, b, dtype=torch.float64)
    ref = P.mc_pr(space, theta, oacq, g["base_samples"], None, grad_estimator="reinforce_ma", ma_counter=3.0,
                  ma_hidden=0.05, grad_output=go)
    pr = product_mc_pr(g, DEV)
    pr._ma_state.copy_(torch.tensor([3.0, 0.05], dtype=torch.float64))
    X = theta.to(DEV).requires_grad_(True)
    val = pr(X)
    grad = torch.autograd.grad(val, X, grad_outputs=go.to(DEV))[0]
    _assert_close(val, ref.value, "PR value")
    _assert_close(grad, ref.grad, "PR gradient", atol=1e-11)
    _assert_close(pr._ma_state, torch.tensor([ref.ma_counter, ref.ma_hidden], dtype=torch.float64), "EMA state",
                  rtol=1e-12)


def test_chunking_is_bitwise_invariant():
    """The L2-resident panel width is a performance knob only: per-column results do not depend on it."""
    g = _problem(500, 128)
    theta = make_theta(6).to(DEV)
    outs = []
    for chunk in (0, 128, 384):
        pr = product_mc_pr(g, DEV)
        pr.acq_function.state.chunk_cols = chunk
        X = theta.clone().requires_grad_(True)
        val = pr(X)
        grad = torch.autograd.grad(val.sum(), X)[0]
        outs.append((val.detach().clone(), grad.clone()))
    # multi-panel calls (the stream-ordered multi-launch sequence) agree bitwise with each other ...
    assert torch.equal(outs[1][0], outs[2][0]) and torch.equal(outs[1][1], outs[2][1])
    # ... and with the single-panel fused step up to the order of the row-tile partial sums
    _assert_close(outs[0][0], outs[1][0], "fused vs multi-panel value", rtol=1e-13, atol=1e-15)
    _assert_close(outs[0][1], outs[1][1], "fused vs multi-panel gradient", rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("n,b,mc,acq", [(1000, 6, 256, "ei"), (300, 5, 128, "ucb")])
def test_fused_step_equals_unfused_sequence(n, b, mc, acq):
    """prb_mc_value_grad on one column panel runs the fused step (prep -> lower -> [epilogue] -> upper -> finish, chained by
    programmatic dependent launch); PRB_OPT_UNFUSED selects the one-kernel-per-stage sequence.  Same arithmetic, different
    kernels and reduction trees: values, gradients, per-sample values and EMA state must agree to rounding."""
    from bo_pr_b200 import _capi as capi
    g = _problem(n, mc, acq)
    theta = make_theta(b).to(DEV)
    res = []
    for unfused in (0, 1):
        pr = product_mc_pr(g, DEV)
        pr.acq_function.state.set_option(capi.PRB_OPT_UNFUSED, unfused)
        pr._ma_state.copy_(torch.tensor([2.0, 0.03], dtype=torch.float64))
        X = theta.clone().requires_grad_(True)
        val = pr(X)
        grad = torch.autograd.grad(val.sum(), X)[0]
        res.append((val.detach().clone(), grad.clone(), pr._ma_state.clone()))
    _assert_close(res[0][0], res[1][0], "value", rtol=1e-13, atol=1e-15)
    _assert_close(res[0][1], res[1][1], "gradient", rtol=1e-11, atol=1e-14)
    _assert_close(res[0][2], res[1][2], "EMA state", rtol=1e-13)


def test_full_size_properties():
    """n = 1000, 64 restarts x 1024 MC samples (the bench configuration): shard invariance (what the multi-GPU split
    relies on), bitwise repeatability, first reinforce_ma backward == plain reinforce, value-only == value of value+grad,
    MC mean consistent with the per-sample values."""
    n, b, mc = 1000, 64, 1024
    g = _problem(n, mc)
    theta = make_theta(b).to(DEV)
    acq = product_acq(g, DEV)
    pr = product_mc_pr(g, DEV, acq=acq)

    def run(th, state=(0.0, 0.0)):
        pr._ma_state.copy_(torch.tensor(state, dtype=torch.float64))
        X = th.clone().requires_grad_(True)
        val = pr(X)
        grad = torch.autograd.grad(val.sum(), X)[0]
        return val.detach().clone(), grad.clone()

    v_all, g_all = run(theta, (5.0, 0.02))
    v_again, g_again = run(theta, (5.0, 0.02))
    assert torch.equal(v_all, v_again) and torch.equal(g_all, g_again)
    # restart shards: ranks evaluate disjoint restart slices with the replicated GP state and the same pre-call EMA state
    for G_ in (2, 8):
        vs, gs = zip(*[run(theta[r::G_], (5.0, 0.02)) for r in range(G_)])
        v_cat = torch.empty_like(v_all)
        g_cat = torch.empty_like(g_all)
        for r in range(G_):
            v_cat[r::G_], g_cat[r::G_] = vs[r], gs[r]
        assert torch.equal(v_cat, v_all) and torch.equal(g_cat, g_all)
    # baseline is 0 while ma_counter == 0
    pr2 = product_mc_pr(dict(g, case=dict(g["case"], grad_estimator="reinforce")), DEV, acq=acq)
    X = theta.clone().requires_grad_(True)
    g_plain = torch.autograd.grad(pr2(X).sum(), X)[0]
    v0, g0 = run(theta, (0.0, 0.0))
    assert torch.equal(g0, g_plain)
    with torch.no_grad():
        pr._ma_state.zero_()
        v_only = pr(theta)
    assert torch.equal(v_only, v0)
    # EMA update: hidden = 0.3 * mean(all acquisition values) after the first call from a zero state
    _assert_close(pr._ma_state[1], 0.3 * v0.mean(), "EMA hidden", rtol=1e-12)
    assert float(pr._ma_state[0]) == 1.0
    assert bool((v0 >= 0).all())  # EI is non-negative


def test_score_function_gradient_is_unbiased_for_the_analytic_objective():
    """on a space small enough to enumerate, the MC score-function gradient converges to the exact
    gradient of the analytic PR objective (both from the CUDA path), and the pathwise continuous gradient likewise."""
    from tests.golden_util import load
    from tests.product_util import product_analytic_pr
    g = load("int_cat_analytic_pm")
    acq = product_acq(g, DEV)
    an = product_analytic_pr(g, DEV, acq=acq)
    theta = g["theta"][3:6].to(DEV)  # interior rows
    X = theta.clone().requires_grad_(True)
    va = an(X)
    ga = torch.autograd.grad(va.sum(), X)[0]
    # independent uniforms (device Philox stream).  NOT the reference's default: it draws the integer and the categorical
    # base samples from two separately scrambled Sobol engines whose i-th points are both scrambles of the same
    # van-der-Corput point, so (u_int[i], u_cat[i]) are dependent and the MC value is biased (0.159 vs 0.172 on row 1
    # here, in the reference's own arithmetic) -- replicated for parity, avoided for this statistical check.
    import bo_pr_b200 as B
    from tests.product_util import integer_bounds
    c = g["case"]
    mc = 1 << 18
    pr = B.MCProbabilisticReparameterization(
        acq_function=acq, dim=c["dim"], integer_indices=c["integer_indices"], integer_bounds=integer_bounds(c, DEV),
        categorical_features=c["categorical_features"], mc_samples=mc, grad_estimator="reinforce", tau=0.1,
        rng="philox", philox_seed=3)
    X2 = theta.clone().requires_grad_(True)
    vm = pr(X2)
    gmc = torch.autograd.grad(vm.sum(), X2)[0]
    assert torch.allclose(vm, va, rtol=0, atol=1e-2)
    # integer columns: the MC estimator omits the Normalize chain-rule factor `range`
    ib = torch.tensor(g["case"]["integer_bounds"], dtype=torch.float64, device=DEV)
    scale = torch.ones(theta.shape[-1], dtype=torch.float64, device=DEV)
    scale[g["case"]["integer_indices"]] = ib[1] - ib[0]
    assert torch.allclose(gmc * scale, ga, rtol=0, atol=5e-2), (gmc * scale - ga).abs().max()


@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 127, 128, 129, 255, 257])
def test_tile_boundary_sizes_vs_oracle(n):
    """Training-set sizes around the 16 / 128 tile boundaries of the triangular products (padding rows, the alpha row
    landing in its own row tile at n = 128, a single training point)."""
    b, mc = 3, 40  # mc not a multiple of the 128-column tile: tiles straddle restarts and the last one is ragged
    g = _problem(n, mc, "ei")
    space, oacq = oracle_space(g), oracle_acq(g)
    theta = make_theta(b, seed=n)
    ref = P.mc_pr(space, theta, oacq, g["base_samples"], None, grad_estimator="reinforce_ma", ma_counter=2.0,
                  ma_hidden=0.01)
    pr = product_mc_pr(g, DEV)
    pr._ma_state.copy_(torch.tensor([2.0, 0.01], dtype=torch.float64))
    X = theta.to(DEV).requires_grad_(True)
    val = pr(X)
    grad = torch.autograd.grad(val.sum(), X)[0]
    _assert_close(val, ref.value, "PR value")
    _assert_close(grad, ref.grad, "PR gradient", atol=1e-11)


def test_empty_and_single_sample_calls():
    g = _problem(64, 1, "ei")
    pr = product_mc_pr(g, DEV)
    space, oacq = oracle_space(g), oracle_acq(g)
    theta = make_theta(2)
    ref = P.mc_pr(space, theta, oacq, g["base_samples"], None, grad_estimator="reinforce_ma")
    X = theta.to(DEV).requires_grad_(True)
    val = pr(X)  # mc = 1
    grad = torch.autograd.grad(val.sum(), X)[0]
    _assert_close(val, ref.value, "PR value (mc = 1)")
    _assert_close(grad, ref.grad, "PR gradient (mc = 1)", atol=1e-11)
    empty = pr(torch.empty(0, 1, 13, dtype=torch.float64, device=DEV))
    assert empty.shape == (0,)


@pytest.mark.parametrize("n,b,mc", [(300, 6, 256), (129, 3, 40), (1000, 16, 1024)])
def test_dedup_equals_plain_path(n, b, mc):
    """PRB_OPT_DEDUP evaluates each unique discrete configuration of a restart once and weights it by its multiplicity:
    same values / gradients / EMA state as the per-sample path up to summation order, and the same as the oracle."""
    g = _problem(n, mc, "ei")
    theta = make_theta(b, seed=3)
    theta[0, 0, 3:] = torch.tensor([0.0, 1.0] * 5, dtype=torch.float64)  # a fully converged restart: one unique config
    acq = product_acq(g, DEV)
    outs = []
    for dd in (False, True):
        pr = product_mc_pr(g, DEV, acq=acq)
        pr.dedup = dd
        pr._ma_state.copy_(torch.tensor([4.0, 0.03], dtype=torch.float64))
        X = theta.to(DEV).requires_grad_(True)
        val = pr(X)
        grad = torch.autograd.grad(val, X, grad_outputs=torch.linspace(0.5, 2.0, b, dtype=torch.float64, device=DEV))[0]
        outs.append((val.detach().clone(), grad.clone(), pr._ma_state.clone()))
    _assert_close(outs[1][0], outs[0][0], "value (dedup vs plain)", rtol=1e-12, atol=1e-15)
    _assert_close(outs[1][1], outs[0][1], "gradient (dedup vs plain)", rtol=1e-11, atol=1e-14)
    _assert_close(outs[1][2], outs[0][2], "EMA state (dedup vs plain)", rtol=1e-12, atol=0)
    if n <= 300:
        space, oacq = oracle_space(g), oracle_acq(g)
        ref = P.mc_pr(space, theta, oacq, g["base_samples"], None, grad_estimator="reinforce_ma", ma_counter=4.0,
                      ma_hidden=0.03, grad_output=torch.linspace(0.5, 2.0, b, dtype=torch.float64))
        _assert_close(outs[1][0], ref.value, "value (dedup vs oracle)")
        _assert_close(outs[1][1], ref.grad, "gradient (dedup vs oracle)", atol=1e-11)


def test_dedup_per_sample_values_and_fused_adam():
    import ctypes as C
    from bo_pr_b200 import _capi as capi
    from bo_pr_b200.acquisition import acq_desc_of
    import bo_pr_b200 as B
    g = _problem(200, 96, "ei")
    acq = product_acq(g, DEV)
    pr = product_mc_pr(g, DEV, acq=acq)
    st = acq.state
    b, mc, d = 5, 96, 13
    th = make_theta(b, seed=9).to(DEV).reshape(b, d).contiguous()
    rnd = pr.input_transform["round"]
    rnd.ensure_base_samples(1, DEV)
    ui, _ = rnd.base_sample_ptrs()
    sp, aq = pr._space(), acq_desc_of(acq)
    res = []
    for dd in (0, 1):
        st.set_option(capi.PRB_OPT_DEDUP, dd)
        value = torch.empty(b, dtype=torch.float64, device=DEV)
        av = torch.empty(b, mc, dtype=torch.float64, device=DEV)
        ws = st.workspace(b * mc, d, b)
        assert st.lib.prb_mc_value_grad(st.handle, C.byref(sp), C.byref(aq), th.data_ptr(), b, mc, ui.data_ptr(), None, 0,
                                        0, 0, None, 0, None, value.data_ptr(), None, av.data_ptr(), ws.data_ptr(),
                                        ws.numel(), 0, None) == 0
        torch.cuda.synchronize()
        res.append((value.clone(), av.clone()))
    assert torch.equal(res[0][1], res[1][1])  # per-sample acquisition values: bitwise identical columns
    _assert_close(res[1][0], res[0][0], "value", rtol=1e-13, atol=0)
    # fused on-device Adam with de-duplication follows the same trajectory
    bounds = torch.stack([torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)]).to(DEV)
    ics = make_theta(4, seed=2).to(DEV)
    outs = []
    for dd in (False, True):
        prx = product_mc_pr(g, DEV, acq=acq)
        prx.dedup = dd
        outs.append(B.gen_candidates_torch(ics, prx, bounds[0], bounds[1], options={"maxiter": 15}))
    assert torch.allclose(outs[0][0], outs[1][0], rtol=0, atol=1e-10)
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-9, atol=1e-13)


def test_svm_shape_53_dims_vs_oracle():
    """the reference's svm / svm_TR configs have 50 binary + 3 continuous parameters (dim 53 > the old 32-column
    limit).  50 binary dims do not fit the 32-bit code of the separable path, so this runs the generic kernels."""
    from tests.random_cases import build_problem
    d = 53
    c = dict(dim=d, integer_indices=list(range(3, d)), integer_bounds=[[0.0] * 50, [1.0] * 50], categorical_features=None,
             kernel_type="mixed", binary_dims=list(range(3, d)), n_train=140, b=4, mc=48, acq="ei", apply_numeric=False,
             grad_estimator="reinforce_ma", seed=41, beta=1.0)
    g = build_problem(c)
    space, oacq = oracle_space(g), oracle_acq(g)
    ref = P.mc_pr(space, g["theta"], oacq, g["base_samples"], None, grad_estimator="reinforce_ma", ma_counter=2.0,
                  ma_hidden=0.01)
    pr = product_mc_pr(g, DEV)
    pr._ma_state.copy_(torch.tensor([2.0, 0.01], dtype=torch.float64))
    X = g["theta"].to(DEV).requires_grad_(True)
    val = pr(X)
    grad = torch.autograd.grad(val.sum(), X)[0]
    _assert_close(val, ref.value, "PR value (d = 53)")
    _assert_close(grad, ref.grad, "PR gradient (d = 53)", atol=1e-11)
    tilde = pr.input_transform(g["theta"].to(DEV).unsqueeze(-3))
    assert torch.equal(tilde.cpu(), ref.tilde_x)
    # multi-start Adam on the same problem: fused device loop == host loop
    import bo_pr_b200 as B
    lo, hi = torch.zeros(d, dtype=torch.float64, device=DEV), torch.ones(d, dtype=torch.float64, device=DEV)
    cf, vf = B.gen_candidates_torch(g["theta"].to(DEV), product_mc_pr(g, DEV), lo, hi, options={"maxiter": 5})
    ch, vh = B.gen_candidates_torch(g["theta"].to(DEV), product_mc_pr(g, DEV), lo, hi, options={"maxiter": 5, "fused": False})
    assert torch.allclose(cf, ch, rtol=0, atol=1e-10) and torch.allclose(vf, vh, rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("kind,n,b,mc", [("rosenbrock", 500, 4, 64), ("rosenbrock", 1000, 3, 32), ("chemistry", 500, 3, 48),
                                         ("chemistry", 1000, 3, 32)])
def test_generic_path_vs_oracle_at_sweep_sizes(kind, n, b, mc):
    """Kernel structures that do not separate (ordinals inside the ARD Matern; mixed_categorical) at several row tiles:
    k* + derivative-factor panel -> lower product -> upper product with the gradient contraction fused into its epilogue.
    Against the oracle, and three ways through the library: fused single-panel step, the multi-launch sequence
    (PRB_OPT_UNFUSED), and narrow column panels (chunk_cols)."""
    from bench_workload import make_generic_problem
    from bo_pr_b200 import _capi as capi
    g, theta = make_generic_problem(kind, n, b, mc, seed=3)
    c = g["case"]
    g["base_samples"] = sobol_base_samples(mc, n_cols=len(c["integer_indices"])) if c["integer_indices"] else None
    g["base_samples_categorical"] = (sobol_base_samples(mc, n_cols=len(c["categorical_features"]), seed=77)
                                     if c["categorical_features"] else None)
    space, oacq = oracle_space(g), oracle_acq(g)
    go = torch.linspace(0.7, 1.3, b, dtype=torch.float64)
    ref = P.mc_pr(space, theta, oacq, g["base_samples"], g["base_samples_categorical"], grad_estimator="reinforce_ma",
                  ma_counter=4.0, ma_hidden=0.2, apply_numeric=c["apply_numeric"], grad_output=go)
    outs = []
    for mode in ("fused", "unfused", "panels"):
        pr = product_mc_pr(g, DEV)
        st = pr.acq_function.state
        st.set_option(capi.PRB_OPT_UNFUSED, int(mode == "unfused"))
        st.chunk_cols = 128 if mode == "panels" else 0
        pr._ma_state.copy_(torch.tensor([4.0, 0.2], dtype=torch.float64))
        X = theta.to(DEV).requires_grad_(True)
        val = pr(X)
        grad = torch.autograd.grad(val, X, grad_outputs=go.to(DEV))[0]
        _assert_close(val, ref.value, f"PR value [{mode}]")
        _assert_close(grad, ref.grad, f"PR gradient [{mode}]", atol=1e-11)
        outs.append((val.detach().clone(), grad.clone()))
    for v, gr in outs[1:]:
        _assert_close(v, outs[0][0], "value across paths", rtol=1e-12, atol=1e-14)
        _assert_close(gr, outs[0][1], "gradient across paths", rtol=1e-10, atol=1e-13)


def test_host_buffer_entry_equals_forward_autograd():
    """prb_mc_value_grad_host (pinned host buffers in and out, one call) == forward + autograd on device tensors, bit for bit,
    EMA state included; also with pageable host memory."""
    g = _problem(300, 128)
    theta = make_theta(6)
    pr_a = product_mc_pr(g, DEV)
    X = theta.to(DEV).requires_grad_(True)
    val = pr_a(X)
    grad = torch.autograd.grad(val.sum(), X)[0]
    for pinned in (True, False):
        pr_b = product_mc_pr(g, DEV)
        th = theta.clone().pin_memory() if pinned else theta.clone()
        v_h = torch.empty(6, dtype=torch.float64)
        g_h = torch.empty(6, 1, 13, dtype=torch.float64)
        if pinned:
            v_h, g_h = v_h.pin_memory(), g_h.pin_memory()
        pr_b.value_and_grad_host(th, v_h, g_h, device=DEV)
        assert torch.equal(v_h, val.detach().cpu()) and torch.equal(g_h, grad.cpu())
        assert torch.equal(pr_a._ma_state, pr_b._ma_state)
    v_only = torch.empty(6, dtype=torch.float64)
    pr_c = product_mc_pr(g, DEV)
    pr_c.value_and_grad_host(theta.clone(), v_only, None, device=DEV)
    assert torch.equal(v_only, val.detach().cpu())
    # gradient rows and values as two views of ONE pinned block ([b, d] then [b]): a single device->host copy returns both
    block = torch.full((6 * 13 + 6,), float("nan"), dtype=torch.float64).pin_memory()
    pr_d = product_mc_pr(g, DEV)
    pr_d.value_and_grad_host(theta.clone().pin_memory(), block[6 * 13:], block[: 6 * 13].view(6, 1, 13), device=DEV)
    assert torch.equal(block[6 * 13:], val.detach().cpu()) and torch.equal(block[: 6 * 13].view(6, 1, 13), grad.cpu())
