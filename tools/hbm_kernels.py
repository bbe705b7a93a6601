"""Achieved HBM bandwidth of the two HBM-bound kernels of the path when their outputs ARE materialised (the parity /
debug entry points; in the fused production path the sampled configurations and the per-sample values never leave the
chip): the PR sampler (prb_pr_sample) and the standalone MC reduction.  Prints JSON; argv[1] = output file."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench_workload import make_problem, make_theta  # noqa: E402
from bo_pr_b200 import _capi as capi  # noqa: E402
from tests.product_util import product_mc_pr  # noqa: E402

PEAK_GBS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                       "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6538.9

dev = torch.device("cuda", 0)
lib = capi.load_library()
g = make_problem(64, seed=0, device=dev)
b, mc, d, n_disc = 2048, 1024, 13, 10
g["case"]["mc"] = mc
g["base_samples"] = g["base_samples_categorical"] = None
pr = product_mc_pr(g, dev)
sp = pr._space()
theta = make_theta(b).to(dev).reshape(b, d).contiguous()
tilde = torch.empty(b, mc, d, dtype=torch.float64, device=dev)
z = torch.empty(b, mc, n_disc, dtype=torch.uint8, device=dev)
prob = torch.empty(b, n_disc, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream()


def run():
    capi.check(lib.prb_pr_sample(C.byref(sp), theta.data_ptr(), b, mc, None, None, 0, 0, tilde.data_ptr(), z.data_ptr(),
                                 prob.data_ptr(), st.cuda_stream), "prb_pr_sample")


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(10):
    run()
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
nbytes = tilde.numel() * 8 + z.numel() + prob.numel() * 8 + theta.numel() * 8
out = {"pr_sample_materialised": {"points": b * mc, "bytes": nbytes, "ms": ms, "gbs": nbytes / ms / 1e6,
                                  "frac_of_measured_hbm": nbytes / ms / 1e6 / PEAK_GBS,
                                  "note": "Philox drawn in-kernel; writes tilde_x [b, mc, 13] f64 + z [b, mc, 10] u8"}}
print(json.dumps(out))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
