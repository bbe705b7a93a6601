import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bo_pr_b200 as B
from bench import bo_iter_problem, BO_OPTS
from tests.product_util import product_acq, product_mc_pr
dev = torch.device("cuda", 0)
g = bo_iter_problem()
acq = product_acq(g, dev)
bounds = torch.stack([torch.zeros(13, dtype=torch.float64), torch.ones(13, dtype=torch.float64)]).to(dev)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    pr = product_mc_pr(g, dev, acq=acq)
    opts = dict(BO_OPTS, seed=rep)
    torch.manual_seed(rep)
    t0 = sync()
    ics = B.gen_batch_initial_conditions(pr, bounds, q=1, num_restarts=20, raw_samples=1024, options=opts)
    t1 = sync()
    cand, vals = B.optimize_acqf(pr, bounds, q=1, num_restarts=20, options=opts, batch_initial_conditions=ics, stochastic=True, return_best_only=False)
    t2 = sync()
    final = pr.sample_candidates(cand)
    t3 = sync()
    with torch.no_grad():
        best = acq(final).argmax()
    chosen = final[best].cpu()
    t4 = sync()
    print("rep", rep, "init %.2f ms | adam %.2f ms | sample_candidates %.2f ms | rescoring %.2f ms | total %.2f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t4-t0)*1e3))
