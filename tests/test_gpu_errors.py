"""Error behaviour of the C ABI on a device: status codes instead of crashes, unsupported configurations raise in the
Python layer instead of silently falling back."""
import ctypes as C

import pytest
import torch

import bo_pr_b200 as B
from bo_pr_b200 import _capi as capi
from bo_pr_b200.acquisition import acq_desc_of
from tests.golden_util import load
from tests.product_util import product_acq, product_mc_pr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_status_codes():
    g = load("ackley13_ei")
    pr = product_mc_pr(g, DEV)
    st = pr.acq_function.state
    lib = st.lib
    sp, aq = pr._space(), acq_desc_of(pr.acq_function)
    b, mc, d = 4, 32, 13
    th = torch.rand(b, d, dtype=torch.float64, device=DEV)
    val = torch.empty(b, dtype=torch.float64, device=DEV)
    ws = st.workspace(b * mc, d, b)

    def call(**kw):
        a = dict(model=st.handle, sp=sp, aq=aq, theta=th.data_ptr(), b=b, mc=mc, est=0, ma=None, upd=0, go=None,
                 val=val.data_ptr(), grad=None, ws=ws.data_ptr(), wsn=ws.numel())
        a.update(kw)
        return lib.prb_mc_value_grad(a["model"], C.byref(a["sp"]), C.byref(a["aq"]), a["theta"], a["b"], a["mc"], None,
                                     None, 0, 0, a["est"], a["ma"], a["upd"], a["go"], a["val"], a["grad"], None, a["ws"],
                                     a["wsn"], 0, None)

    assert call() == 0
    assert call(wsn=1024) == -3                       # PRB_ERR_WORKSPACE
    assert call(est=7) == -1                          # unknown estimator
    assert call(est=1) == -1                          # reinforce_ma without the EMA state buffer
    assert call(go=th.data_ptr()) == -1               # gradient requested without an output buffer
    assert call(mc=0) == -1
    bad = capi.PrbSpaceDesc.from_buffer_copy(bytes(sp))
    bad.d = 12                                        # one-hot / kernel-input dimension mismatch with the model
    bad.int_cols[9] = 11
    assert call(sp=bad) == -1
    bad_aq = capi.PrbAcqDesc.from_buffer_copy(bytes(aq))
    bad_aq.kind = 9
    assert call(aq=bad_aq) == -1
    torch.cuda.synchronize()
    assert b"workspace" in lib.prb_status_string(-3)


def test_model_create_rejects_bad_descriptors():
    lib = capi.load_library()
    desc = capi.PrbModelDesc()
    h = C.c_void_p()
    assert lib.prb_model_create(C.byref(h), C.byref(desc), None) == -1          # n = 0, NULL pointers
    X = torch.rand(8, 3, dtype=torch.float64, device=DEV)
    L = torch.eye(8, dtype=torch.float64, device=DEV)
    al = torch.zeros(8, dtype=torch.float64, device=DEV)
    desc.n, desc.d_k, desc.X_train, desc.Linv, desc.alpha = 8, 3, X.data_ptr(), L.data_ptr(), al.data_ptr()
    desc.n_terms = 3
    assert lib.prb_model_create(C.byref(h), C.byref(desc), None) == -2          # more additive terms than the ABI holds
    desc.n_terms = 1
    desc.terms[0].n_factors = 1
    desc.terms[0].outputscale = 1.0
    desc.terms[0].factors[0].kind = 0
    desc.terms[0].factors[0].n_dims = 1
    desc.terms[0].factors[0].power = 1
    desc.terms[0].factors[0].dims[0] = 5                                        # column outside d_k
    desc.terms[0].factors[0].lengthscale[0] = 1.0
    assert lib.prb_model_create(C.byref(h), C.byref(desc), None) == -1
    desc.terms[0].factors[0].dims[0] = 2
    desc.terms[0].factors[0].lengthscale[0] = -1.0                              # non-positive lengthscale
    assert lib.prb_model_create(C.byref(h), C.byref(desc), None) == -1
    desc.terms[0].factors[0].lengthscale[0] = 0.7
    assert lib.prb_model_create(C.byref(h), C.byref(desc), None) == 0
    assert lib.prb_model_n(h) == 8 and lib.prb_model_dk(h) == 3
    assert lib.prb_model_set_option(h, 99, 1) == -1
    assert lib.prb_model_destroy(h) == 0


def test_python_layer_refuses_unsupported_inputs():
    g = load("ackley13_ei")
    acq = product_acq(g, DEV)
    pr = product_mc_pr(g, DEV, acq=acq)
    with pytest.raises(capi.PrbError):
        pr(torch.rand(2, 1, 13, dtype=torch.float64))                 # CPU tensor: no CPU path
    with pytest.raises(NotImplementedError):
        pr(torch.rand(2, 2, 13, dtype=torch.float64, device=DEV))     # q > 1
    with pytest.raises(ValueError):
        pr(torch.rand(2, 1, 12, dtype=torch.float64, device=DEV))     # wrong dimension
    with pytest.raises(NotImplementedError):
        B.MCProbabilisticReparameterization(acq, dim=13, integer_indices=list(range(3, 13)),
                                            integer_bounds=torch.tensor([[0.] * 10, [1.] * 10], device=DEV),
                                            grad_estimator="arm")
    with pytest.raises(NotImplementedError):
        B.ExpectedImprovement(acq.model, best_f=0.0, posterior_transform=object())
    with pytest.raises(AssertionError):
        acq(torch.rand(2, 3, 13, dtype=torch.float64, device=DEV))    # analytic AFs need q = 1

    class FakeTanimoto:  # a kernel type libprb200 has no device code for
        pass
    from bo_pr_b200.kernels import kernel_terms
    with pytest.raises(NotImplementedError):
        kernel_terms(FakeTanimoto(), 3)


def test_gp_state_is_rebuilt_when_hyperparameters_change():
    """the cached device GP state must not outlive the hyper-parameters it was built from.  Evaluate, assign a
    new lengthscale / mutate the targets in place, evaluate again: the result equals a freshly built model's."""
    from tests.golden_util import load
    from tests.product_util import product_acq, product_model
    g = load("ackley13_ei")
    model = product_model(g, DEV)
    acq = product_acq(g, DEV, model=model)
    X = g["theta"][:4].to(DEV)
    with torch.no_grad():
        v0 = acq(X).clone()
    leaf = model.covar_module.base_kernel.kernels[0]
    new_ls = (leaf.lengthscale * 1.7).clone()
    leaf.lengthscale = new_ls  # no refresh()
    with torch.no_grad():
        v1 = acq(X).clone()
    g2 = dict(g, hp=dict(g["hp"], cont_lengthscale=new_ls.reshape(-1).tolist()))
    with torch.no_grad():
        v_fresh = product_acq(g2, DEV)(X)
    assert not torch.allclose(v0, v1)
    assert torch.equal(v1, v_fresh)
    model.train_targets.mul_(0.5)  # in-place mutation bumps the tensor's version counter
    with torch.no_grad():
        v2 = acq(X)
    assert not torch.allclose(v1, v2)
