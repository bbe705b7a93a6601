"""Per-CTA phase trace of the two triangular products of one bench-workload step (needs the PRB_TRACE_CTAS build):

    PRB_LIB_NAME=libprb200_trace.so PRB_NVCC_EXTRA=-DPRB_TRACE_CTAS=1 python -m bo_pr_b200.build --force
    PRB_LIB_NAME=libprb200_trace.so python tools/debug/cta_trace.py

For every SM: the share of the kernel's span during which 0 / 1 / 2 of its resident CTAs are inside their k loop (between
the first full-barrier wait and the barrier after the loop), and the mean prologue / k-loop / epilogue length in cycles."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench_workload import make_problem, make_theta  # noqa: E402
from bo_pr_b200 import _capi as capi  # noqa: E402
from bo_pr_b200.acquisition import acq_desc_of  # noqa: E402
from tests.product_util import product_acq, product_mc_pr  # noqa: E402

dev = torch.device("cuda", 0)
lib = capi.load_library()
n, b, mc = 1000, 64, 1024
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
                                     capi.PRB_EST_REINFORCE_MA, ma.data_ptr(), 1, ones.data_ptr(), value.data_ptr(),
                                     gradt.data_ptr(), None, ws.data_ptr(), ws.numel(), 0, st.cuda_stream), "step")


for _ in range(3):
    step()
torch.cuda.synchronize()
STRIDE = 8 * 16384
buf = torch.zeros(4 * STRIDE, dtype=torch.int64, device=dev)
lib.prb_debug_set_cta_trace.argtypes = [C.c_void_p]
lib.prb_debug_set_cta_trace.restype = None
lib.prb_debug_set_cta_trace(buf.data_ptr())
step()
torch.cuda.synchronize()
lib.prb_debug_set_cta_trace(None)
tr = buf.cpu().numpy().reshape(4, 16384, 8)

out = {}
for li, name in ((0, "lower"), (1, "upper")):
    t = tr[li]
    t = t[t[:, 1] != 0]
    res = {"ctas": int(len(t))}
    pro, loop, epi = t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3]
    res["prologue_cycles_mean"] = float(pro.mean())
    res["loop_cycles_mean"] = float(loop.mean())
    res["epilogue_cycles_mean"] = float(epi.mean())
    cov = np.zeros(3)
    span_total = 0.0
    resident = np.zeros(3)
    for sm in np.unique(t[:, 0]):
        e = t[t[:, 0] == sm]
        lo, hi = e[:, 1].min(), e[:, 4].max()
        span_total += hi - lo
        for arr, a_col, b_col in ((cov, 2, 3), (resident, 1, 4)):
            ev = sorted([(x, 1) for x in e[:, a_col]] + [(x, -1) for x in e[:, b_col]])
            cur, last = 0, lo
            for x, dlt in ev:
                arr[min(cur, 2)] += x - last
                last = x
                cur += dlt
            arr[0] += hi - last
    res["in_loop_share_0_1_2"] = [float(c / span_total) for c in cov]
    res["resident_share_0_1_2"] = [float(c / span_total) for c in resident]
    res["span_cycles_mean_per_sm"] = float(span_total / len(np.unique(t[:, 0])))
    out[name] = res
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/cta_trace.json", "w"), indent=1)
np.save("gpurun_out/cta_trace.npy", tr[:2, :8192, :5])
