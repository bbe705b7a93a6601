"""Synthetic PR-EI sweep of BASELINE.json configs[4]: n_train 200-4000, 64 restarts x {128, 1024} MC samples, 1 GPU.
Prints one JSON object per shape (CUDA-event timing of prb_mc_value_grad, device-resident inputs) and writes them to the
file given as argv[1].  Roofline fraction is against F_vg(n) = 2 n^2 + 100 n flop per evaluation
and the measured FP64 DMMA peak (37.1 TFLOP/s)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench_workload import make_problem, make_theta  # noqa: E402
from bo_pr_b200 import _capi as capi  # noqa: E402
from bo_pr_b200.acquisition import acq_desc_of  # noqa: E402
from tests.product_util import product_acq, product_mc_pr  # noqa: E402

PEAK = 37.1e12


def measure(n, b, mc, steps=5, warmup=2, grad=True):
    dev = torch.device("cuda", 0)
    lib = capi.load_library()
    g = make_problem(n, seed=0, device=dev)
    g["case"]["mc"] = mc
    g["base_samples"] = g["base_samples_categorical"] = None
    acq = product_acq(g, dev)
    state = acq.state
    pr = product_mc_pr(dict(g, case=dict(g["case"], grad_estimator="reinforce_ma")), dev, acq=acq)
    theta = make_theta(b).to(dev).reshape(b, -1).contiguous()
    d = theta.shape[-1]
    sp, aq = pr._space(), acq_desc_of(acq)
    ones = torch.ones(b, dtype=torch.float64, device=dev)
    value = torch.empty(b, dtype=torch.float64, device=dev)
    gradt = torch.empty(b, d, dtype=torch.float64, device=dev)
    ma = torch.zeros(2, dtype=torch.float64, device=dev)
    ws = state.workspace(b * mc, d, b)
    st = torch.cuda.current_stream()

    def step():
        capi.check(lib.prb_mc_value_grad(state.handle, C.byref(sp), C.byref(aq), theta.data_ptr(), b, mc, None, None, 0, 0,
                                         capi.PRB_EST_REINFORCE_MA, ma.data_ptr(), 1, ones.data_ptr() if grad else None,
                                         value.data_ptr(), gradt.data_ptr() if grad else None, None, ws.data_ptr(),
                                         ws.numel(), 0, st.cuda_stream), "prb_mc_value_grad")

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    evals = b * mc / (ms * 1e-3)
    flop = (2.0 * n * n + 100.0 * n) if grad else (1.0 * n * n + 70.0 * n)
    return {"n_train": n, "restarts": b, "mc": mc, "mode": "value+grad" if grad else "value", "ms_per_step": ms,
            "evals_per_s": evals, "roofline_frac": evals * flop / PEAK, "workspace_mb": ws.numel() / 2 ** 20}


if __name__ == "__main__":
    out = []
    for n in (200, 500, 1000, 2000, 4000):
        for mc in (128, 1024):
            r = measure(n, 64, mc)
            print(json.dumps(r), flush=True)
            out.append(r)
    r = measure(1000, 64, 1024, grad=False)
    print(json.dumps(r), flush=True)
    out.append(r)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
