"""GP-state builder on the device: K(X, X) -> blocked Cholesky -> triangular inverse -> alpha ->
packed factors, all in libprb200 (csrc/prb_chol.cu).  Checked against the oracle's factorisation (CPU, float64), against the
defining identities at sizes the CPU oracle would take long for, and against the cuSOLVER-based cross-check builder."""
import pytest
import torch

from bench_workload import make_problem
from bo_pr_b200 import _capi as capi
from bo_pr_b200.kernels import kernel_terms
from bo_pr_b200.models import GPState
from oracle import gp_oracle as G
from tests.golden_util import load, oracle_spec
from tests.product_util import close, product_covar

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _state(g, builder):
    Xk, y = g["train_X_model"].to(DEV), g["train_Y"].to(DEV)
    covar = product_covar(g["case"], g["hp"], Xk, y)
    return GPState(Xk, y, kernel_terms(covar, Xk.shape[-1]), g["noise"], g["mean_const"], builder=builder)


@pytest.mark.parametrize("n", [1, 2, 37, 63, 64, 65, 119, 128, 129, 257, 1000])
def test_factor_and_alpha_vs_oracle(n):
    """Tile-boundary sizes (one block, block + 1, power-of-two padding) up to the benchmark size."""
    g = make_problem(n, seed=2)
    st = _state(g, "device")
    ogp = G.OracleGP(g["train_X_model"], g["train_Y"], oracle_spec(g), g["noise"], g["mean_const"])
    scale = float(ogp.Linv.abs().max())
    err = float((st.Linv.cpu() - ogp.Linv).abs().max())
    assert err <= 1e-9 * scale, f"Linv: max abs error {err:.3g} vs scale {scale:.3g}"
    assert float(st.Linv.cpu().triu(1).abs().max()) == 0.0 if n > 1 else True
    ok, worst = close(st.alpha, ogp.alpha, 1e-8, 1e-9 * float(ogp.alpha.abs().max()))
    assert ok, f"alpha: {worst:.3g} x tolerance"


@pytest.mark.parametrize("name", ["chemistry_mc_ei", "mixed_int_f1_ei", "named_chemistry_analytic_ei"])
def test_builders_agree_on_posteriors(name):
    """Device builder vs the cuSOLVER cross-check builder through the hot path: posterior mean / variance at points near and
    far from the training data."""
    g = load(name)
    gen = torch.Generator().manual_seed(1)
    Xk = g["train_X_model"]
    idx = torch.randint(0, Xk.shape[0], (64,), generator=gen)
    Xt = Xk[idx].clone()
    cont = g["cont_dims"]
    Xt[:, cont] = (Xt[:, cont] + 0.1 * torch.randn(64, len(cont), generator=gen, dtype=torch.float64)).clamp(0, 1)
    Xt[:8, cont] = Xk[idx[:8]][:, cont] + 1e-3
    desc = capi.PrbAcqDesc(kind=capi.PRB_ACQ_MEAN, ei_variant=0, param=0.0, min_variance=1e-10, acq_var_floor=0.0)
    outs = []
    for builder in ("device", "torch"):
        st = _state(g, builder)
        _, mean, var, _ = st.acq_points(desc, Xt.to(DEV), want_post=True)
        outs.append((mean.clone(), var.clone()))
    ok, w = close(outs[0][0], outs[1][0], 1e-9, 1e-12)
    assert ok, f"posterior mean: {w:.3g} x tolerance"
    ok, w = close(outs[0][1], outs[1][1], 1e-9, 1e-12)
    assert ok, f"posterior variance: {w:.3g} x tolerance"


def test_large_factorisation_identities():
    """n = 4000 (the top of BASELINE.json's sweep): Linv (K + noise) Linv^T = I and (K + noise) alpha = y - c, evaluated on the
    device with plain torch matmuls (test-side arithmetic only)."""
    n = 4000
    g = make_problem(n, seed=0, device=DEV)
    st = _state(g, "device")
    K = torch.empty(n, n, dtype=torch.float64, device=DEV)
    assert st.lib.prb_kernel_matrix(__import__("ctypes").byref(st._desc), st.X_train.data_ptr(), n, K.data_ptr(), None) == 0
    torch.cuda.synchronize()
    K.diagonal().add_(st.noise)
    R = st.Linv @ K @ st.Linv.T - torch.eye(n, dtype=torch.float64, device=DEV)
    assert float(R.abs().max()) <= 1e-9, float(R.abs().max())
    res = K @ st.alpha - (st.train_Y - st.mean_const)
    assert float(res.abs().max()) <= 1e-8 * float(st.train_Y.abs().max())
    assert float(st.Linv.triu(1).abs().max()) == 0.0


def test_not_positive_definite_is_reported():
    g = make_problem(40, seed=1)
    g = dict(g, noise=-2.0)  # K - 2 I has negative eigenvalues (k(x, x) = 1.3): a non-positive pivot must be reported
    with pytest.raises(capi.PrbError, match="positive definite"):
        _state(g, "device")
