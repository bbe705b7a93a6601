"""ctypes binding of ``include/prb200.h`` (the C ABI of ``libprb200.so``).

There is no CPU implementation behind these calls.  Importing this module works without a GPU (so the package can
be imported and the ABI inspected on a build box); *calling* anything that computes requires the shared library and
an sm_100 device, and fails loudly otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

PRB_MAX_DIMS = 64
PRB_MAX_FACTORS = 4
PRB_MAX_TERMS = 2
PRB_MAX_CAT = 8
PRB_MAX_CARD = 32
PRB_MAX_MODELS = 8
PRB_MACQ_CEI = 0
PRB_MACQ_EHVI = 1

PRB_FACTOR_MATERN52 = 0
PRB_FACTOR_CATEGORICAL = 1
PRB_ACQ_EI, PRB_ACQ_UCB, PRB_ACQ_MEAN = 0, 1, 2
PRB_EST_REINFORCE, PRB_EST_REINFORCE_MA = 0, 1
PRB_OPT_DEDUP = 1
PRB_OPT_UNFUSED = 2
PRB_PROFILE_SLOTS = 12
PROFILE_SLOT_NAMES = ["pr_sample", "kstar_panel", "trgemm_lower", "acq_epilogue", "trgemm_upper", "grad_contract",
                      "sum_splits", "mc_reduce", "ma_update", "analytic", "adam", "other"]

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", os.environ.get("PRB_LIB_NAME", "libprb200.so"))

EXPORTS = [
    "prb_version", "prb_abi_sizeof", "prb_debug_small_timestamps", "prb_debug_fp64_peak", "prb_launch_count", "prb_profile_enable", "prb_profile_read", "prb_status_string", "prb_last_cuda_error", "prb_model_create", "prb_model_create_from_data", "prb_rff_model_create", "prb_model_destroy", "prb_model_set_option",
    "prb_model_n", "prb_model_dk", "prb_workspace_bytes", "prb_philox_uniform", "prb_pr_sample",
    "prb_analytic_enumerate", "prb_kernel_matrix", "prb_acq_points", "prb_sample_candidates", "prb_finalize_candidates",
    "prb_trust_region_bounds", "prb_workspace_bytes_multi", "prb_mc_value_grad_multi", "prb_acq_points_multi",
    "prb_mc_value_grad", "prb_mc_value_grad_host", "prb_mc_screen", "prb_analytic_value_grad", "prb_mc_adam", "prb_analytic_adam",
]


class PrbFactor(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_dims", C.c_int32), ("power", C.c_int32), ("reserved", C.c_int32),
                ("dims", C.c_int32 * PRB_MAX_DIMS), ("lengthscale", C.c_double * PRB_MAX_DIMS)]


class PrbTerm(C.Structure):
    _fields_ = [("outputscale", C.c_double), ("additive", C.c_int32), ("n_factors", C.c_int32),
                ("factors", PrbFactor * PRB_MAX_FACTORS)]


class PrbModelDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("d_k", C.c_int32), ("X_train", C.c_void_p), ("Linv", C.c_void_p),
                ("alpha", C.c_void_p), ("mean_const", C.c_double), ("n_terms", C.c_int32), ("reserved", C.c_int32),
                ("terms", PrbTerm * PRB_MAX_TERMS)]


class PrbSpaceDesc(C.Structure):
    _fields_ = [("d", C.c_int32), ("n_int", C.c_int32), ("n_cat", C.c_int32), ("apply_numeric", C.c_int32),
                ("int_cols", C.c_int32 * PRB_MAX_DIMS), ("int_lo", C.c_double * PRB_MAX_DIMS),
                ("int_hi", C.c_double * PRB_MAX_DIMS), ("cat_start", C.c_int32 * PRB_MAX_CAT),
                ("cat_card", C.c_int32 * PRB_MAX_CAT), ("tau", C.c_double), ("raw_integer_space", C.c_int32),
                ("reserved", C.c_int32), ("latent_dim", C.c_int32 * PRB_MAX_CAT), ("latent_tables", C.c_void_p)]


class PrbAcqDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ei_variant", C.c_int32), ("param", C.c_double), ("min_variance", C.c_double),
                ("acq_var_floor", C.c_double)]


class PrbMultiAcqDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_models", C.c_int32), ("objective_index", C.c_int32), ("reserved", C.c_int32),
                ("best_f", C.c_double), ("min_variance", C.c_double), ("sigma_floor", C.c_double),
                ("has_lower", C.c_int32 * PRB_MAX_MODELS), ("has_upper", C.c_int32 * PRB_MAX_MODELS),
                ("lower", C.c_double * PRB_MAX_MODELS), ("upper", C.c_double * PRB_MAX_MODELS),
                ("var_floor", C.c_double), ("n_cells", C.c_int32), ("reserved2", C.c_int32), ("cell_bounds", C.c_void_p)]


class PrbError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """dlopen ``libprb200.so`` and declare the prototypes.  Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PrbError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  bo_pr_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double, C.c_size_t
    lib.prb_version.restype = C.c_int
    lib.prb_launch_count.restype = C.c_uint64
    lib.prb_launch_count.argtypes = []
    lib.prb_abi_sizeof.argtypes = [C.c_int]
    lib.prb_debug_small_timestamps.argtypes = [vp]
    lib.prb_debug_fp64_peak.argtypes = [C.c_int, vp, vp]
    lib.prb_profile_enable.argtypes = [C.c_int]
    lib.prb_profile_read.argtypes = [vp, vp, i32]
    lib.prb_status_string.restype = C.c_char_p
    lib.prb_status_string.argtypes = [C.c_int]
    lib.prb_last_cuda_error.restype = C.c_char_p
    lib.prb_model_create.argtypes = [C.POINTER(vp), C.POINTER(PrbModelDesc), vp]
    lib.prb_model_create_from_data.argtypes = [C.POINTER(vp), C.POINTER(PrbModelDesc), vp, vp, vp, vp, vp]
    lib.prb_rff_model_create.argtypes = [C.POINTER(vp), C.c_int32, C.c_int32, vp, vp, vp, C.c_double, vp]
    lib.prb_model_destroy.argtypes = [vp]
    lib.prb_model_set_option.argtypes = [vp, i32, i64]
    lib.prb_model_n.argtypes = [vp]
    lib.prb_model_dk.argtypes = [vp]
    lib.prb_workspace_bytes.restype = sz
    lib.prb_workspace_bytes.argtypes = [vp, i64, i32, i32, i32]
    lib.prb_philox_uniform.argtypes = [u64, u64, i32, i32, vp, vp]
    lib.prb_pr_sample.argtypes = [C.POINTER(PrbSpaceDesc), vp, i32, i32, vp, vp, u64, u64, vp, vp, vp, vp]
    lib.prb_sample_candidates.argtypes = [C.POINTER(PrbSpaceDesc), vp, i32, vp, u64, u64, vp, vp, vp]
    lib.prb_finalize_candidates.argtypes = [vp, C.POINTER(PrbSpaceDesc), C.POINTER(PrbAcqDesc), vp, i32, vp, u64, u64, vp,
                                            vp, vp, vp, vp, sz, vp]
    lib.prb_trust_region_bounds.argtypes = [vp, vp, i32, i32, i32, dbl, vp, vp, vp, vp, vp, vp]
    lib.prb_workspace_bytes_multi.restype = sz
    lib.prb_workspace_bytes_multi.argtypes = [C.POINTER(vp), i32, i64, i32, i32]
    lib.prb_mc_value_grad_multi.argtypes = [C.POINTER(vp), i32, C.POINTER(PrbSpaceDesc), C.POINTER(PrbMultiAcqDesc), vp, i32,
                                            i32, vp, vp, u64, u64, i32, vp, i32, vp, vp, vp, vp, vp, sz, vp]
    lib.prb_acq_points_multi.argtypes = [C.POINTER(vp), i32, C.POINTER(PrbMultiAcqDesc), vp, i64, vp, vp, vp, vp, sz, vp]
    lib.prb_analytic_enumerate.argtypes = [C.POINTER(PrbSpaceDesc), vp, i32, vp, i32, vp, vp, vp]
    lib.prb_kernel_matrix.argtypes = [C.POINTER(PrbModelDesc), vp, i64, vp, vp]
    lib.prb_acq_points.argtypes = [vp, C.POINTER(PrbAcqDesc), vp, i64, vp, vp, vp, vp, vp, sz, i32, vp]
    lib.prb_mc_value_grad.argtypes = [vp, C.POINTER(PrbSpaceDesc), C.POINTER(PrbAcqDesc), vp, i32, i32, vp, vp, u64,
                                      u64, i32, vp, i32, vp, vp, vp, vp, vp, sz, i32, vp]
    lib.prb_mc_value_grad_host.argtypes = [vp, C.POINTER(PrbSpaceDesc), C.POINTER(PrbAcqDesc), vp, i32, i32, vp, vp, u64,
                                           u64, i32, vp, i32, i32, vp, vp, vp, sz, i32, vp]
    lib.prb_mc_screen.argtypes = [vp, C.POINTER(PrbSpaceDesc), C.POINTER(PrbAcqDesc), vp, i32, i32, i32, vp, vp, u64, u64,
                                  i32, vp, vp, vp, sz, i32, vp]
    lib.prb_analytic_value_grad.argtypes = [vp, C.POINTER(PrbSpaceDesc), C.POINTER(PrbAcqDesc), vp, i32, vp, i32, vp,
                                            vp, vp, vp, vp, vp, sz, i32, vp]
    lib.prb_mc_adam.argtypes = [vp, C.POINTER(PrbSpaceDesc), C.POINTER(PrbAcqDesc), vp, i32, i32, vp, vp, u64, u64,
                                i32, vp, vp, vp, dbl, i32, i32, vp, vp, vp, sz, i32, vp]
    lib.prb_analytic_adam.argtypes = [vp, C.POINTER(PrbSpaceDesc), C.POINTER(PrbAcqDesc), vp, i32, vp, i32, vp, vp, dbl, i32,
                                      i32, vp, vp, vp, sz, i32, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("prb_version",):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status == 0:
        return
    lib = load_library()
    msg = lib.prb_status_string(status).decode()
    if status in (-4, -5, -6):
        msg += ": " + lib.prb_last_cuda_error().decode()
    raise PrbError(f"{what} failed: {msg} (status {status})")


def require_cuda(t, name: str = "tensor"):
    """Every tensor that crosses the ABI must be a contiguous CUDA tensor; there is no host path."""
    if not t.is_cuda:
        raise PrbError(f"{name} must live on a CUDA device: bo_pr_b200 has no CPU implementation (got {t.device}).")
    return t


def ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()
