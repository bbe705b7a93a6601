"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and exports every function that
include/prb200.h declares; the ctypes mirrors of the descriptor structs have the C layout; compute entry points refuse
to run without a device instead of falling back."""
import ctypes as C
import os
import re

import pytest
import torch

import __graft_entry__ as entry
from bo_pr_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return capi.load_library()


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "prb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(prb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared_functions()
    assert len(names) >= 18
    for nm in names:
        assert hasattr(lib, nm), f"libprb200.so does not export {nm}"
    assert sorted(capi.EXPORTS) == names, "bo_pr_b200/_capi.py EXPORTS is out of sync with include/prb200.h"


def test_version_and_status_strings(lib):
    assert lib.prb_version() == 200
    assert lib.prb_status_string(0) == b"ok"
    assert b"workspace" in lib.prb_status_string(-3)


def test_struct_layout_matches_c(lib):
    for which, ty in enumerate([capi.PrbFactor, capi.PrbTerm, capi.PrbModelDesc, capi.PrbSpaceDesc, capi.PrbAcqDesc,
                                capi.PrbMultiAcqDesc]):
        assert lib.prb_abi_sizeof(which) == C.sizeof(ty), ty.__name__


def test_argument_validation_without_device(lib):
    sp = capi.PrbSpaceDesc()
    sp.d, sp.n_int, sp.n_cat, sp.tau = 3, 0, 0, 0.1  # no discrete parameter at all
    assert lib.prb_pr_sample(C.byref(sp), None, 1, 1, None, None, 0, 0, None, None, None, None) == -1
    sp.d = 1000
    assert lib.prb_pr_sample(C.byref(sp), None, 1, 1, None, None, 0, 0, None, None, None, None) == -2
    assert lib.prb_workspace_bytes(None, 10, 1, 3, 0) == 0
    assert lib.prb_mc_value_grad(None, None, None, None, 1, 1, None, None, 0, 0, 0, None, 0, None, None, None, None, None,
                                 0, 0, None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import bo_pr_b200 as B
    X = torch.rand(4, 3, dtype=torch.float64)
    with pytest.raises(capi.PrbError):
        B.FixedNoiseGP(X, torch.rand(4, 1, dtype=torch.float64), torch.full((4, 1), 1e-6, dtype=torch.float64),
                       covar_module=B.get_kernel("mixed", 3, [2], {}, X, None))
    tf = B.MCProbabilisticReparameterizationInputTransform(integer_indices=[2], integer_bounds=torch.tensor([[0.], [1.]]))
    with pytest.raises(capi.PrbError):
        tf.transform(X.reshape(4, 1, 1, 3))


def test_host_side_reference_error_behaviour():
    """Error surface of the reference constructors (SURVEY.md section 8b 'Errors')."""
    import bo_pr_b200 as B
    with pytest.raises(NotImplementedError):
        B.MCProbabilisticReparameterization(acq_function=B.AcquisitionFunction(), dim=3)
    with pytest.raises(ValueError):
        B.MCProbabilisticReparameterizationInputTransform()
    with pytest.raises(ValueError):
        B.get_kernel("nope", 3, [], {}, None, None)
