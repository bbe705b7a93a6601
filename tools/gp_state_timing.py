"""GP-state build time per BO iteration: the library's own builder (prb_model_create_from_data:
K assembly, blocked Cholesky, divide-and-conquer triangular inverse, alpha, packing -- csrc/prb_chol.cu) next to the
cuSOLVER / cuBLAS cross-check path (torch.linalg.cholesky + solve_triangular + cholesky_solve + prb_model_create)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench_workload import make_problem  # noqa: E402
from bo_pr_b200.kernels import kernel_terms  # noqa: E402
from bo_pr_b200.models import GPState  # noqa: E402
from tests.product_util import product_covar  # noqa: E402

dev = torch.device("cuda", 0)
out = []
for n in (119, 500, 1000, 2000, 4000):
    g = make_problem(n, seed=0, device=dev)
    Xk, y = g["train_X_model"].to(dev), g["train_Y"].to(dev)
    terms = kernel_terms(product_covar(g["case"], g["hp"], Xk, y), Xk.shape[-1])
    row = {"n_train": n}
    for builder in ("device", "torch"):
        ts = []
        for rep in range(6):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st = GPState(Xk, y, terms, g["noise"], g["mean_const"], builder=builder)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
            del st
        row[f"ms_{builder}"] = sorted(ts[1:])[len(ts[1:]) // 2] * 1e3
    print(json.dumps(row), flush=True)
    out.append(row)
if len(sys.argv) > 1:
    json.dump({"what": "wall-clock ms per GP-state build (median of 5 after one warm-up), device builder (prb_chol.cu) vs "
                       "torch.linalg (cuSOLVER potrf + cuBLAS trsm) + prb_model_create", "rows": out},
              open(sys.argv[1], "w"), indent=1)
