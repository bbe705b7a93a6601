"""Where does the host time of one end-to-end step (bench.py's e2e leg) go?  cProfile over the Python path + wall-clock split."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from bench_workload import make_problem, make_theta
from tests.product_util import product_acq, product_mc_pr
dev = torch.device("cuda", 0)
n, b, mc = 1000, (int(sys.argv[1]) if len(sys.argv) > 1 else 64), 1024
problem = make_problem(n, seed=0); problem["case"]["mc"] = mc; problem["base_samples"] = None; problem["base_samples_categorical"] = None
acq = product_acq(problem, dev)
pr = product_mc_pr(dict(problem, case=dict(problem["case"], grad_estimator="reinforce_ma")), dev, acq=acq)
theta = make_theta(b, seed=1)
theta_host = theta.clone().pin_memory(); val_host = torch.empty(b, dtype=torch.float64).pin_memory(); grad_host = torch.empty(b, 1, theta.shape[-1], dtype=torch.float64).pin_memory()
T = {}
def e2e_step(prof=False):
    t0 = time.perf_counter()
    X = theta_host.to(dev, non_blocking=True).requires_grad_(True)
    t1 = time.perf_counter()
    v = pr(X)
    t2 = time.perf_counter()
    g = torch.autograd.grad(v.sum(), X)[0]
    t3 = time.perf_counter()
    val_host.copy_(v.detach(), non_blocking=True); grad_host.copy_(g, non_blocking=True)
    t4 = time.perf_counter()
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    for k, d in zip(("h2d", "forward_enqueue", "backward_enqueue", "d2h_enqueue", "sync_wait", "total"), (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)):
        T[k] = T.get(k, 0.0) + d
for _ in range(5): e2e_step()
T.clear()
N = 30
for _ in range(N): e2e_step()
print({k: round(v / N * 1e6, 1) for k, v in T.items()}, "us per step")
pr_ = cProfile.Profile(); pr_.enable()
for _ in range(N): e2e_step()
pr_.disable()
pstats.Stats(pr_).sort_stats("cumulative").print_stats(28)

# ---- split of the forward enqueue: Python before the library call / the library call itself / Python after
cc = pr._call_ctx(dev)
real_fn = cc["fn"]
TT = {"c": 0.0, "n": 0}
def timed_fn(*a):
    t = time.perf_counter(); r = real_fn(*a); TT["c"] += time.perf_counter() - t; TT["n"] += 1; return r
cc["fn"] = timed_fn
T.clear()
for _ in range(N): e2e_step()
print("library call (prb_mc_value_grad, 5 launches):", round(TT["c"] / TT["n"] * 1e6, 1), "us of forward_enqueue", round(T["forward_enqueue"] / N * 1e6, 1), "us")
cc["fn"] = real_fn
X = theta_host.to(dev).requires_grad_(True)
for name, f in (("module __call__ + Function.apply + _launch (no grad)", lambda: pr(X.detach())), ("_launch only", lambda: pr._launch(X, True)),
                ("gp_state_from_model", lambda: __import__("bo_pr_b200").models.gp_state_from_model(pr.acq_function.model)),
                ("acq_desc_of", lambda: __import__("bo_pr_b200").acquisition.acq_desc_of(pr.acq_function)),
                ("torch.empty x2", lambda: (torch.empty(b, dtype=torch.float64, device=dev), torch.empty(b, 13, dtype=torch.float64, device=dev)))):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(200): f()
    dt = (time.perf_counter() - t) / 200; torch.cuda.synchronize()
    print(f"{name}: {dt * 1e6:.1f} us")
