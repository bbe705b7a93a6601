"""Time of the step before the hot path: building the device GP state of one BO iteration -- K(X, X) by
libprb200's kernel, Cholesky / triangular inverse / alpha by torch (cuSOLVER, cuBLAS), prb_model_create (packing, TMA
descriptors, separable-structure tables).  Prints one JSON object per n and writes them to argv[1]."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench_workload import make_problem  # noqa: E402
from tests.product_util import product_model  # noqa: E402

dev = torch.device("cuda", 0)
out = []
for n in (119, 500, 1000, 2000, 4000):
    g = make_problem(n, seed=0, device=dev)
    times = []
    for rep in range(4):
        model = product_model(g, dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = model.state  # builds GPState
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        del model, st
    r = {"n_train": n, "gp_state_build_ms": sorted(times[1:])[1] * 1e3,
         "what": "prb_kernel_matrix + torch.linalg.cholesky + solve_triangular(L, I) + cholesky_solve + prb_model_create"}
    out.append(r)
    print(json.dumps(r), flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
