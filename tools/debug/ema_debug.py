import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from bench_workload import make_problem, make_theta, sobol_base_samples
from oracle import pr_oracle as P
from tests.golden_util import oracle_acq, oracle_space
from tests.product_util import product_mc_pr
torch.set_printoptions(precision=17)
n, b, mc = 200, 8, 128
g = make_problem(n, seed=0)
g["case"].update(acq="ei", mc=mc, grad_estimator="reinforce_ma")
g["base_samples"] = sobol_base_samples(mc); g["base_samples_categorical"] = None
space, oacq = oracle_space(g), oracle_acq(g)
theta = make_theta(b)
ref = P.mc_pr(space, theta, oacq, g["base_samples"], None, grad_estimator="reinforce_ma", ma_counter=3.0, ma_hidden=0.05)
pr = product_mc_pr(g, "cuda:0")
pr._ma_state.copy_(torch.tensor([3.0, 0.05], dtype=torch.float64))
X = theta.to("cuda:0").requires_grad_(True)
val = pr(X)
print("val", val.detach().cpu()); print("ref", ref.value)
print("mean val", val.detach().cpu().mean().item(), "mean ref", ref.value.mean().item(), "mean acq", ref.acq_values.mean().item())
print("state", pr._ma_state.cpu(), "ref", ref.ma_counter, ref.ma_hidden)
print("expected", 0.05 - (0.05 - val.detach().cpu().mean().item()) * (1 - 0.7))
