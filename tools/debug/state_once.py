import sys, os
sys.path.insert(0, '/root/repo')
import torch
from bench_workload import make_problem
from bo_pr_b200.kernels import kernel_terms
from bo_pr_b200.models import GPState
from tests.product_util import product_covar
dev = torch.device("cuda", 0)
n = int(sys.argv[1])
g = make_problem(n, seed=0, device=dev)
Xk, y = g["train_X_model"].to(dev), g["train_Y"].to(dev)
terms = kernel_terms(product_covar(g["case"], g["hp"], Xk, y), Xk.shape[-1])
for _ in range(2):
    st = GPState(Xk, y, terms, g["noise"], g["mean_const"], builder="device")
    torch.cuda.synchronize()
