"""Candidate finalisation on the device: the tail of a BO iteration after the multi-start optimisation.

``discrete_mixed_bo/run_one_replication.py:500-528``: draw one discrete configuration per restart from p(theta*)
(``sample_candidates``), map one-hot blocks to numeric codes when the model wants them, score the sampled rows with the BASE
acquisition function, keep the best restart.  Here that is one call, ``prb_finalize_candidates`` -- sampler kernel (Philox
stream or explicit uniforms), posterior + acquisition kernels on the b sampled rows, argmax kernel -- with no host round
trip between the stages; the caller reads back one row.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _capi as capi
from .acquisition import acq_desc_of
from .models import gp_state_from_model


def finalize_candidates(pr, candidates: Tensor, seed: Optional[int] = None, offset: int = 1 << 32,
                        u: Optional[Tensor] = None, return_all: bool = False):
    """``candidates [b, 1, d]`` (the optimiser's continuous PR parameters) -> ``(best [1, d], best value)`` and, with
    ``return_all``, also ``(sampled [b, 1, d], values [b])``.  ``pr`` is an MC or analytic PR acquisition function; the
    scoring uses ``pr.acq_function`` (the wrapped EI / UCB / posterior mean), exactly as the reference re-scores with
    ``true_acq_func``.  Draws: Philox ``(seed, offset)`` with counter (restart, parameter), or explicit ``u``."""
    capi.require_cuda(candidates, "candidates")
    if candidates.ndim != 3 or candidates.shape[-2] != 1:
        raise NotImplementedError("expected candidates of shape b x q=1 x d")
    b, d = candidates.shape[0], candidates.shape[-1]
    dev = candidates.device
    if hasattr(pr.acq_function, "multi_desc"):
        # model lists: sampler kernel, then the model-list acquisition on the sampled rows; argmax with device ops
        rows, xk = pr._sample_candidates_device(candidates, 0 if seed is None else int(seed), int(offset), u, want_xk=True)
        with torch.no_grad():
            vals = pr.acq_function(xk.unsqueeze(-2))
        i = torch.argmax(vals)
        if return_all:
            return rows[i], vals[i], rows, vals
        return rows[i], vals[i]
    state = gp_state_from_model(pr.acq_function.model)
    sp, aq = pr._space(), acq_desc_of(pr.acq_function)
    theta = candidates.detach().to(torch.float64).reshape(b, d).contiguous()
    out_x = torch.empty(b, d, dtype=torch.float64, device=dev)
    out_a = torch.empty(b, dtype=torch.float64, device=dev)
    best = torch.empty(1 + d, dtype=torch.float64, device=dev)
    if u is not None:
        u = capi.require_cuda(u, "u").detach().to(torch.float64).reshape(b, -1).contiguous()
        if u.shape[1] != sp.n_int + sp.n_cat:
            raise ValueError(f"u must have {sp.n_int + sp.n_cat} columns")
    ws = state.workspace(b, state.d_k)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        capi.check(state.lib.prb_finalize_candidates(state.handle, C.byref(sp), C.byref(aq), theta.data_ptr(), b,
                                                     capi.ptr(u), 0 if seed is None else int(seed), int(offset),
                                                     out_x.data_ptr(), out_a.data_ptr(), best.data_ptr(), None,
                                                     ws.data_ptr(), ws.numel(), st), "prb_finalize_candidates")
    chosen, value = best[1:].reshape(1, d).to(candidates.dtype), best[0]
    if return_all:
        return chosen, value, out_x.reshape(b, 1, d).to(candidates.dtype), out_a
    return chosen, value
