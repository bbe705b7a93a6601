import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bo_pr_b200 as B
from tests.config_timings import CONFIGS, OPTS, bounds_for, ts_gp, DEV
from tests.product_util import product_mc_pr
from tests.random_cases import build_problem
c = dict(CONFIGS["ackley13 pr__ts"], b=20, mc=128, seed=3)
g = build_problem(c); gp = ts_gp(g); bounds = bounds_for(c).to(DEV)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(4):
    torch.manual_seed(rep)
    t0 = sync()
    sm = B.get_gp_samples(gp, num_outputs=1, n_samples=1, num_rff_features=512)
    t1 = sync()
    acq = B.PosteriorMean(sm); pr = product_mc_pr(g, DEV, acq=acq); _ = sm.state
    t2 = sync()
    Xr = torch.rand(1024, 1, 13, dtype=torch.float64, device=DEV)
    y = pr.screen(Xr, 32)
    t3 = sync()
    cand, vals = B.gen_candidates_torch(Xr[:5].clone(), pr, bounds[0], bounds[1], options={"maxiter": 200})
    t4 = sync()
    cand, vals = B.optimize_acqf(pr, bounds, q=1, num_restarts=20, raw_samples=1024, options=dict(OPTS, seed=rep, nonnegative=False), stochastic=True, return_best_only=False)
    t5 = sync()
    print("draw sample %.2f ms | state+pr %.2f | screen 1024x128 %.2f | one batch 5x200 adam %.2f | optimize_acqf %.2f" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)))
