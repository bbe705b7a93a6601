"""Per-BO-iteration acquisition-optimisation time for the four named experiment configs of BASELINE.json (shapes of
SURVEY.md App. D, synthetic GP data at the end-of-run training-set size), CUDA path vs the CPU oracle port.

    python tests/config_timings.py [out.json] [--no-cpu]   (uses the oracle as the timed CPU baseline: lives under tests/)

ackley13 pr__ei (MC, binary dims: separable path, persistent kernel), rosenbrock10_scaled pr__ucb (MC, ordinals inside the
ARD Matern: generic path), mixed_int_f1_TR pr__ei (MC, ordinals + 2 binary dims, trust-region bounds: generic path),
chemistry_analytic pr__ei (analytic enumeration of 192 configurations, mixed_categorical kernel)."""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bo_pr_b200 as B  # noqa: E402
from oracle import pr_oracle as P  # noqa: E402
from tests.golden_util import oracle_acq, oracle_space  # noqa: E402
from tests.product_util import product_acq, product_analytic_pr, product_mc_pr  # noqa: E402
from tests.random_cases import build_problem  # noqa: E402

DEV = torch.device("cuda", 0)
OPTS = {"batch_limit": 5, "init_batch_limit": 32, "maxiter": 200}

CONFIGS = {
    "ackley13 pr__ei": dict(dim=13, integer_indices=list(range(3, 13)), integer_bounds=[[0.0] * 10, [1.0] * 10],
                            categorical_features=None, kernel_type="mixed", binary_dims=list(range(3, 13)), n_train=119,
                            acq="ei", apply_numeric=False, grad_estimator="reinforce_ma", analytic=False),
    "rosenbrock10_scaled pr__ucb": dict(dim=10, integer_indices=list(range(4, 10)), integer_bounds=[[0.0] * 6, [3.0] * 6],
                                        categorical_features=None, kernel_type="mixed", binary_dims=[], n_train=119,
                                        acq="ucb", beta=0.2 * 10 * math.log(2 * 100), apply_numeric=False,
                                        grad_estimator="reinforce_ma", analytic=False),
    "mixed_int_f1_TR pr__ei": dict(dim=16, integer_indices=list(range(8, 16)),
                                   integer_bounds=[[0.0] * 8, [1.0, 1.0, 2.0, 2.0, 4.0, 4.0, 6.0, 6.0]],
                                   categorical_features=None, kernel_type="mixed", binary_dims=[8, 9], n_train=119, acq="ei",
                                   apply_numeric=False, grad_estimator="reinforce_ma", analytic=False, trust_region=True),
    # Thompson sampling over an RFF posterior sample (rffs.py, experiment_utils.py:481-492): BoTorch's feature map only
    # accepts Scale(Matern), so the GP here uses one ARD Matern over all 13 inputs; the sample draw (m x m Cholesky, once per
    # BO iteration) is inside the timed segment
    "ackley13 pr__ts": dict(dim=13, integer_indices=list(range(3, 13)), integer_bounds=[[0.0] * 10, [1.0] * 10],
                            categorical_features=None, kernel_type="mixed", binary_dims=list(range(3, 13)), n_train=119,
                            acq="pm", apply_numeric=False, grad_estimator="reinforce_ma", analytic=False, ts=True),
    "chemistry_analytic pr__ei": dict(dim=22, integer_indices=[], integer_bounds=[[], []],
                                      categorical_features={2: 4, 3: 12, 4: 4}, kernel_type="mixed_categorical",
                                      binary_dims=[], n_train=219, acq="ei", apply_numeric=True, grad_estimator=None,
                                      analytic=True),
}


def bounds_for(c):
    d = c["dim"]
    lo, hi = torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)
    if c.get("trust_region"):  # run_one_replication.py:410-440 at the initial TurboState (length 0.8)
        from bo_pr_b200.trust_region import TurboState, trust_region_bounds
        centre = torch.full((1, d), 0.45, dtype=torch.float64)
        return trust_region_bounds(TurboState(dim=d, batch_size=1, is_constrained=False), centre,
                                   torch.zeros(1, 1, dtype=torch.float64), torch.stack([lo, hi]),
                                   binary_dims=c["binary_dims"])
    return torch.stack([lo, hi])


def ts_gp(g):
    """Scale(Matern-5/2 ARD over every input) GP on the config's training data (what RandomFourierFeatures accepts)."""
    from bo_pr_b200.kernels import MaternKernel, ScaleKernel
    X, y = g["train_X_model"].to(DEV), g["train_Y"].to(DEV)
    base = MaternKernel(nu=2.5, ard_num_dims=X.shape[-1])
    base.lengthscale = torch.full((1, X.shape[-1]), 0.7, dtype=torch.float64)
    covar = ScaleKernel(base)
    covar.outputscale = torch.tensor(1.0, dtype=torch.float64)
    return B.FixedNoiseGP(X, y.unsqueeze(-1), torch.full_like(y, 1e-4).unsqueeze(-1), covar_module=covar)


def ours(c, g, repeats=3):
    gp = ts_gp(g) if c.get("ts") else None
    acq = None if c.get("ts") else product_acq(g, DEV)
    bounds = bounds_for(c).to(DEV)
    times = []
    for rep in range(repeats + 1):
        opts = dict(OPTS, seed=rep, nonnegative=(c["acq"] == "ei"))
        torch.manual_seed(rep)
        if not c.get("ts"):
            pr = product_analytic_pr(g, DEV, acq=acq) if c["analytic"] else product_mc_pr(g, DEV, acq=acq)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if c.get("ts"):  # a fresh posterior sample per BO iteration, drawn inside the timed segment
            acq = B.PosteriorMean(B.get_gp_samples(gp, num_outputs=1, n_samples=1, num_rff_features=512))
            pr = product_mc_pr(g, DEV, acq=acq)
        cand, vals = B.optimize_acqf(pr, bounds, q=1, num_restarts=20, raw_samples=1024, options=opts, stochastic=True,
                                     return_best_only=False)
        chosen, _ = B.finalize_candidates(pr, cand, seed=rep)  # device: sample + numeric map + base-AF re-score + argmax
        chosen.cpu()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    return sorted(times[1:])[len(times[1:]) // 2]


def cpu(c, g):
    from bo_pr_b200.initializers import draw_sobol_samples, initialize_q_batch, initialize_q_batch_nonneg
    torch.set_num_threads(os.cpu_count() or 1)
    space = oracle_space(g)
    if c.get("ts"):  # the same kind of sample, evaluated by the CPU oracle (drawn on the GPU box's device, copied over)
        from oracle import gp_oracle as G
        from oracle.rff_oracle import OracleRFFSample
        torch.manual_seed(0)
        sm = B.get_gp_samples(ts_gp(g), num_outputs=1, n_samples=1, num_rff_features=512)
        oacq = G.OraclePosteriorMean(OracleRFFSample(sm.basis.weights, sm.basis.bias, sm.weights, sm.basis.scale))
    else:
        oacq = oracle_acq(g)
    st = {"c": 0.0, "h": 0.0}

    def vg(X, with_grad=True):
        if c["analytic"]:
            v, gr, _, _ = P.analytic_pr(space, X, oacq, apply_numeric=c["apply_numeric"], with_grad=with_grad)
            return v, gr
        r = P.mc_pr(space, X, oacq, g["base_samples"], g["base_samples_categorical"], grad_estimator=c["grad_estimator"],
                    ma_counter=st["c"], ma_hidden=st["h"], apply_numeric=c["apply_numeric"], with_grad=with_grad)
        st["c"], st["h"] = r.ma_counter, r.ma_hidden
        return r.value, r.grad

    bounds = bounds_for(c)
    torch.manual_seed(0)
    t0 = time.perf_counter()
    X_rnd = draw_sobol_samples(bounds, 1024, 1, seed=0)
    Y = torch.cat([vg(X_rnd[s:s + 32], False)[0] for s in range(0, 1024, 32)])
    ics = (initialize_q_batch_nonneg if c["acq"] == "ei" else initialize_q_batch)(X_rnd, Y, 20)
    for s in range(0, 20, 5):
        P.adam_multistart(lambda X: vg(X, True), lambda X: vg(X, False)[0], ics[s:s + 5], bounds[0], bounds[1], maxiter=200)
    return time.perf_counter() - t0


if __name__ == "__main__":
    out = {}
    do_cpu = "--no-cpu" not in sys.argv
    for name, c in CONFIGS.items():
        c = dict(c, b=20, mc=128, seed=3)
        g = build_problem(c)
        r = {"sec_ours": ours(c, g)}
        if do_cpu:
            r["sec_cpu_port"] = cpu(c, g)
            r["cpu_cores"] = torch.get_num_threads()
        out[name] = r
        print(name, json.dumps(r), flush=True)
    files = [a for a in sys.argv[1:] if not a.startswith("--")]
    if files:
        json.dump({"what": "acquisition-optimisation segment per BO iteration: 1024 raw samples screened in chunks of 32, "
                           "20 restarts x 200 Adam steps in batches of 5, candidate sampling, base-AF re-scoring; "
                           "synthetic GP data at end-of-run n_train", "configs": out}, open(files[0], "w"), indent=1)
