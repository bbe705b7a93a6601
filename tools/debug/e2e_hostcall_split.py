"""Host-side split of the host-buffer entry (bench.py's e2e step): enqueue-only time of value_and_grad_host, the library call
inside it, and the synchronous step."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from bench_workload import make_problem, make_theta  # noqa: E402
from bo_pr_b200 import _capi as capi  # noqa: E402
from tests.product_util import product_acq, product_mc_pr  # noqa: E402

dev = torch.device("cuda", 0)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
problem = make_problem(1000, seed=0)
problem["case"]["mc"] = 1024
problem["base_samples"] = problem["base_samples_categorical"] = None
acq = product_acq(problem, dev)
pr = product_mc_pr(dict(problem, case=dict(problem["case"], grad_estimator="reinforce_ma")), dev, acq=acq)
theta = make_theta(b, seed=1)
d = theta.shape[-1]
th = theta.clone().pin_memory()
blk = torch.empty(b * d + b, dtype=torch.float64).pin_memory()
gh, vh = blk[: b * d].view(b, 1, d), blk[b * d:]
for _ in range(10):
    pr.value_and_grad_host(th, vh, gh)
N = 200
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(N):
    pr.value_and_grad_host(th, vh, gh, synchronize=True)
sync_step = (time.perf_counter() - t) / N
lib = capi.load_library()
real = lib.prb_mc_value_grad_host
acc = {"t": 0.0}


def timed(*a):
    t0 = time.perf_counter()
    r = real(*a)
    acc["t"] += time.perf_counter() - t0
    return r


torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(N):
    pr.value_and_grad_host(th, vh, gh, synchronize=False)
enq = (time.perf_counter() - t) / N
torch.cuda.synchronize()
lib.prb_mc_value_grad_host = timed
for _ in range(N):
    pr.value_and_grad_host(th, vh, gh, synchronize=False)
torch.cuda.synchronize()
lib.prb_mc_value_grad_host = real
print(f"b={b}: synchronous step {sync_step * 1e6:.1f} us | enqueue-only call {enq * 1e6:.1f} us (GPU-bound if ~ step time) | "
      f"library call inside it {acc['t'] / N * 1e6:.1f} us")
# pure Python overhead: call with b = 0 rows is a no-op; instead time the pieces
from bo_pr_b200.models import gp_state_from_model  # noqa: E402
from bo_pr_b200.acquisition import acq_desc_of  # noqa: E402
for name, f in (("gp_state_from_model", lambda: gp_state_from_model(pr.acq_function.model)),
                ("acq_desc_of", lambda: acq_desc_of(pr.acq_function)),
                ("state.workspace", lambda: acq.state.workspace(b * 1024, d, b)),
                ("set_option", lambda: acq.state.set_option(capi.PRB_OPT_DEDUP, 0))):
    t = time.perf_counter()
    for _ in range(2000):
        f()
    print(f"  {name}: {(time.perf_counter() - t) / 2000 * 1e6:.2f} us")
