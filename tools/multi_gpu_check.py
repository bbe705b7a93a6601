"""torchrun check of the sharded multi-start optimisation (one process per GPU, NCCL): every rank ends with the same
candidates and values as a single-process run over all restarts with the same per-restart batching.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bo_pr_b200 as B  # noqa: E402
from bench_workload import make_problem, make_theta, sobol_base_samples  # noqa: E402
from tests.product_util import product_mc_pr  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = make_problem(150, seed=2)
g["case"].update(mc=128, grad_estimator="reinforce")  # stateless estimator: sharding cannot change the trajectory
g["base_samples"] = sobol_base_samples(128)
g["base_samples_categorical"] = None
bounds = torch.stack([torch.zeros(13, dtype=torch.float64), torch.ones(13, dtype=torch.float64)]).to(dev)
ics = make_theta(8, seed=4).to(dev)
opts = {"maxiter": 30, "batch_limit": 1}  # one restart per batch on every rank and in the single-process run
pr = product_mc_pr(g, dev)
cand_s, val_s = B.optimize_acqf(pr, bounds, q=1, num_restarts=8, options=dict(opts), batch_initial_conditions=ics,
                                stochastic=True, return_best_only=False, shard=True)
pr1 = product_mc_pr(g, dev)
cand_1, val_1 = B.optimize_acqf(pr1, bounds, q=1, num_restarts=8, options=dict(opts), batch_initial_conditions=ics,
                                stochastic=True, return_best_only=False, shard=False)
ok = torch.equal(cand_s, cand_1) and torch.equal(val_s, val_1)
best_s = B.optimize_acqf(pr, bounds, q=1, num_restarts=8, options=dict(opts), batch_initial_conditions=ics,
                         stochastic=True, return_best_only=True, shard=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("sharded == single-process:", bool(flag.item()), "| best value", float(best_s[1]), "world", world)
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
