"""Generate golden vectors by running the REFERENCE's own L1 code (imported unmodified from /root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (it needs /root/reference, which does not exist on the
GPU box):

    python oracle/gen_golden.py            # writes tests/golden/*.pt

What runs: ``discrete_mixed_bo.probabilistic_reparameterization.{MCProbabilisticReparameterization,
AnalyticProbabilisticReparameterization}`` and the input transforms of ``discrete_mixed_bo/input.py``, through the
BoTorch-surface stand-in in ``oracle/shim`` (BoTorch/GPyTorch are absent from the image), wrapping the L0 oracle
``oracle/gp_oracle.py`` as the base acquisition function.  Each case stores the seeded inputs (GP data and
hyperparameters, theta, base samples) and the reference's outputs: PR value, gradient, per-sample acquisition
values, the sampled configurations ``tilde_x``, the rounding components ``z``, the probabilities ``p`` and the EMA
baseline state after each call.

Cases follow the shapes of BASELINE.json's configs at small n so the fixtures stay small.
"""
from __future__ import annotations

import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import gp_oracle as G  # noqa: E402

from discrete_mixed_bo.input import LatentCategoricalEmbedding, LatentCategoricalSpec  # noqa: E402  (reference code)
from discrete_mixed_bo.probabilistic_reparameterization import (  # noqa: E402  (reference code)
    AnalyticProbabilisticReparameterization,
    MCProbabilisticReparameterization,
)

torch.set_default_dtype(torch.float64)


def make_cases():
    """Static description of every golden case (shared with tests through the saved files)."""
    cases = {}
    # A. ackley13 pr__ei : 3 continuous + 10 binary, kernel "mixed" (ARD Matern x iso Matern on binaries)
    cases["ackley13_ei"] = dict(
        dim=13, integer_indices=list(range(3, 13)), integer_bounds=[[0.0] * 10, [1.0] * 10], categorical_features=None,
        kernel_type="mixed", binary_dims=list(range(3, 13)), n_train=40, b=6, mc=32, acq="ei", apply_numeric=False,
        grad_estimator="reinforce_ma", analytic=False, calls=3, seed=11,
    )
    # B. rosenbrock10_scaled pr__ucb : 4 continuous + 6 ordinals {0..3}, ARD Matern over all dims
    cases["rosenbrock10_ucb"] = dict(
        dim=10, integer_indices=list(range(4, 10)), integer_bounds=[[0.0] * 6, [3.0] * 6], categorical_features=None,
        kernel_type="mixed", binary_dims=[], n_train=30, b=7, mc=32, acq="ucb", beta=0.2 * 10 * math.log(2 * 5),
        apply_numeric=False, grad_estimator="reinforce_ma", analytic=False, calls=2, seed=12, edge_rows=True,
    )
    # C. chemistry : 2 continuous + categoricals (4, 12, 4), mixed_categorical kernel on the numeric layout
    chem = dict(
        dim=22, integer_indices=[], integer_bounds=[[], []], categorical_features={2: 4, 3: 12, 4: 4},
        kernel_type="mixed_categorical", binary_dims=[], n_train=30, b=4, acq="ei", apply_numeric=True, seed=13,
    )
    cases["chemistry_mc_ei"] = dict(chem, mc=64, grad_estimator="reinforce", analytic=False, calls=2)
    cases["chemistry_analytic_ei"] = dict(chem, mc=0, grad_estimator=None, analytic=True, calls=1)
    # C'. the same search-space shape under kernel_type "mixed_latent": the model embeds the numeric codes with the
    # reference's LatentCategoricalEmbedding (latent dims 2, 2, 1 as experiment_utils.py:297-303 picks them)
    lat = dict(chem, dim=21, categorical_features={2: 4, 3: 12, 4: 3}, kernel_type="mixed_latent", seed=17)
    cases["chemistry_latent_mc_ei"] = dict(lat, mc=64, grad_estimator="reinforce_ma", analytic=False, calls=2)
    cases["chemistry_latent_analytic_ei"] = dict(lat, mc=0, grad_estimator=None, analytic=True, calls=1, n_train=25, b=3)
    # D. mixed_int_f1 : 8 continuous + 8 integers with ranges 1,1,2,2,4,4,6,6 (two binary dims)
    cases["mixed_int_f1_ei"] = dict(
        dim=16, integer_indices=list(range(8, 16)),
        integer_bounds=[[0.0] * 8, [1.0, 1.0, 2.0, 2.0, 4.0, 4.0, 6.0, 6.0]], categorical_features=None,
        kernel_type="mixed", binary_dims=[8, 9], n_train=30, b=5, mc=32, acq="ei", apply_numeric=False,
        grad_estimator="reinforce_ma", analytic=False, calls=2, seed=14,
    )
    # E. integers with negative bounds + one-hot categoricals under the plain "mixed" kernel (no numeric map)
    cases["neg_int_onehot_ei"] = dict(
        dim=9, integer_indices=[2, 3], integer_bounds=[[-2.0, 0.0], [2.0, 5.0]], categorical_features={4: 3, 5: 2},
        kernel_type="mixed", binary_dims=[], n_train=25, b=6, mc=48, acq="ei", apply_numeric=False,
        grad_estimator="reinforce", analytic=False, calls=1, seed=15, edge_rows=True,
    )
    # F. small analytic case with integers (probabilities that lose mass at the upper bound) + a categorical
    cases["int_cat_analytic_pm"] = dict(
        dim=5, integer_indices=[1, 2], integer_bounds=[[0.0, 0.0], [3.0, 1.0]], categorical_features={3: 2},
        kernel_type="mixed", binary_dims=[2], n_train=20, b=6, mc=0, acq="pm", apply_numeric=False,
        grad_estimator=None, analytic=True, calls=1, seed=16, edge_rows=True,
    )
    # ---- the same four shapes at the sizes the named configs actually run at: end-of-run n_train,
    # mc = 128, restart batches of 5 (one case with all 20 restarts).  One call each to keep the fixtures small.
    cases["named_ackley13_ei"] = dict(cases["ackley13_ei"], n_train=119, b=20, mc=128, calls=1, seed=21)
    cases["named_rosenbrock10_ucb"] = dict(cases["rosenbrock10_ucb"], n_train=119, b=5, mc=128, calls=1, seed=22,
                                           beta=0.2 * 10 * math.log(2 * 100))
    cases["named_chemistry_analytic_ei"] = dict(cases["chemistry_analytic_ei"], n_train=219, b=5, seed=23)
    cases["named_mixed_int_f1_ei"] = dict(cases["mixed_int_f1_ei"], n_train=119, b=5, mc=128, calls=1, seed=24)
    return cases


def build_problem(c):
    """Seeded GP data + hyperparameters for a case.  Returns dict of tensors / plain values."""
    g = torch.Generator().manual_seed(c["seed"])
    dim = c["dim"]
    cf = c["categorical_features"] or {}
    ints = c["integer_indices"]
    ib = torch.tensor(c["integer_bounds"], dtype=torch.float64).reshape(2, -1)
    n = c["n_train"]
    X = torch.rand(n, dim, generator=g)
    # integers: valid grid points, normalised
    for k, col in enumerate(ints):
        lo, hi = ib[0, k], ib[1, k]
        v = torch.floor(lo + X[:, col] * (hi - lo + 1)).clamp(lo, hi)
        X[:, col] = (v - lo) / (hi - lo)
    # categoricals: valid one-hot rows
    if cf:
        start = min(cf.keys())
        s = start
        for _, card in cf.items():
            cat = torch.randint(0, card, (n,), generator=g)
            X[:, s:s + card] = torch.nn.functional.one_hot(cat, card).to(X)
            s += card
    # what the model sees
    if c["apply_numeric"]:
        start = min(cf.keys())
        cols = [X[:, :start]]
        s = start
        for _, card in cf.items():
            cols.append(X[:, s:s + card].argmax(-1, keepdim=True).to(X))
            s += card
        Xk = torch.cat(cols, dim=-1)
    else:
        Xk = X
    latent = None
    X_numeric = Xk
    ctf = cf if c["kernel_type"] == "mixed_categorical" else {}
    if c["kernel_type"] == "mixed_latent":
        # experiment_utils.py:289-313: one LatentCategoricalSpec per numeric categorical column, latent dim 1 (card <= 3)
        # or 2; categorical_transformed_features = {first latent column: latent dim}
        start = min(cf.keys())
        specs, ctf, col = [], {}, start
        for j, (idx, card) in enumerate(cf.items()):
            ldim = 1 if card <= 3 else 2
            specs.append(LatentCategoricalSpec(idx=start + j, num_categories=card, latent_dim=ldim))
            ctf[col] = ldim
            col += ldim
        emb = LatentCategoricalEmbedding(specs, dim=X_numeric.shape[-1])
        emb.eval()
        with torch.no_grad():
            for sp_ in specs:  # stand-in for fitted tables
                getattr(emb, f"raw_latent_emb_tables_{sp_.idx}").copy_(
                    0.8 * torch.randn(sp_.num_categories, sp_.latent_dim, generator=g))
            Xk = emb.transform(X_numeric)
            tables = [emb.get_emb_table(sp_.idx).detach().clone() for sp_ in specs]
        latent = dict(specs=[(sp_.idx, sp_.num_categories, sp_.latent_dim) for sp_ in specs], tables=tables, ctf=ctf,
                      transform=emb.transform)
    dk = Xk.shape[-1]
    binary = c["binary_dims"]
    if c["kernel_type"] == "mixed_categorical":
        cat_dims = list(cf.keys())
    elif c["kernel_type"] == "mixed_latent":
        cat_dims = list(range(min(ctf.keys()), dk))
    else:
        cat_dims = []
    cont_dims = sorted(set(range(dk)) - set(binary) - set(cat_dims))
    hp = dict(
        cont_lengthscale=(0.3 + 0.9 * torch.rand(len(cont_dims), generator=g)).tolist(),
        binary_lengthscale=[1.2] if binary else [],
        categorical_lengthscale=(0.5 + 1.5 * torch.rand(len(cat_dims), generator=g)).tolist()
        if c["kernel_type"] == "mixed_categorical" else [],
        outputscale=1.3, outputscale_sum=0.4,
    )
    if latent is not None:
        hp["latent_lengthscale"] = (0.6 + 1.2 * torch.rand(len(ctf), generator=g)).tolist()
    spec = G.reference_kernel_spec(c["kernel_type"], dk, binary, ctf, **hp)
    # y: a draw from the GP prior, standardised
    K = G.eval_kernel(spec, Xk, Xk) + 1e-6 * torch.eye(n)
    y = torch.linalg.cholesky(K) @ torch.randn(n, generator=g)
    y = (y - y.mean()) / y.std()
    noise = 1e-6
    mean_const = 0.1
    return dict(X_onehot=X, X_model=Xk, X_numeric=X_numeric, y=y, hp=hp, noise=noise, mean_const=mean_const, spec=spec,
                cont_dims=cont_dims, cat_dims=cat_dims, latent=latent)


def make_theta(c, g):
    b, dim = c["b"], c["dim"]
    theta = torch.rand(b, 1, dim, generator=g)
    if c.get("edge_rows"):
        # rows that hit the clamp / (x == 1) / exact-integer branches
        theta[0, 0, c["integer_indices"]] = 1.0
        theta[1, 0, c["integer_indices"]] = 0.0
        if b > 2 and len(c["integer_indices"]) > 0:
            theta[2, 0, c["integer_indices"][0]] = 0.5
    return theta


def make_acq(c, prob):
    lat = prob.get("latent")
    # mixed_latent: the model-level input transform is the REFERENCE's LatentCategoricalEmbedding.transform
    model = G.OracleGP(prob["X_model"], prob["y"], prob["spec"], prob["noise"], prob["mean_const"],
                       input_transform=None if lat is None else lat["transform"])
    if c["acq"] == "ei":
        return G.OracleEI(model, best_f=prob["y"].max())
    if c["acq"] == "ucb":
        return G.OracleUCB(model, beta=c["beta"])
    return G.OraclePosteriorMean(model)


def run_case(name, c):
    prob = build_problem(c)
    g = torch.Generator().manual_seed(c["seed"] + 1000)
    theta = make_theta(c, g)
    acq = make_acq(c, prob)
    ib = torch.tensor(c["integer_bounds"], dtype=torch.float64).reshape(2, -1)
    out = dict(case=dict(c), theta=theta, train_X_onehot=prob["X_onehot"], train_X_model=prob["X_model"],
               train_Y=prob["y"], hp=prob["hp"], noise=prob["noise"], mean_const=prob["mean_const"],
               cont_dims=prob["cont_dims"], cat_dims=prob["cat_dims"], calls=[])
    if prob.get("latent") is not None:
        out["train_X_numeric"] = prob["X_numeric"]
        out["latent"] = {k: v for k, v in prob["latent"].items() if k != "transform"}
    if c["analytic"]:
        pr = AnalyticProbabilisticReparameterization(
            acq_function=acq, dim=c["dim"], dtype=torch.float64, device=torch.device("cpu"),
            integer_indices=c["integer_indices"],
            integer_bounds=ib, categorical_features=c["categorical_features"], apply_numeric=c["apply_numeric"], tau=0.1)
        X = theta.clone().requires_grad_(True)
        val = pr(X)
        grad = torch.autograd.grad(val.sum(), X)[0]
        rnd = pr.input_transform["round"]
        if "unnormalize" in pr.input_transform:
            xun = pr.input_transform["unnormalize"](theta)
        else:
            xun = theta
        out["all_discrete_options"] = rnd.all_discrete_options.clone()
        out["calls"].append(dict(value=val.detach(), grad=grad, probs=rnd.get_probs(xun).squeeze(-1).detach()))
    else:
        torch.manual_seed(c["seed"] + 7)  # seeds the Sobol seeds drawn by torch.randint inside transform()
        pr = MCProbabilisticReparameterization(
            acq_function=acq, dim=c["dim"],
            integer_indices=c["integer_indices"],
            integer_bounds=ib, categorical_features=c["categorical_features"], batch_limit=16,
            apply_numeric=c["apply_numeric"], mc_samples=c["mc"], grad_estimator=c["grad_estimator"], tau=0.1)
        for call in range(c["calls"]):
            X = (theta if call == 0 else (theta + 0.01 * call).clamp(0, 1)).clone().requires_grad_(True)
            pre = (float(pr.ma_counter), float(pr.ma_hidden))
            val = pr(X)
            go = torch.linspace(0.5, 1.5, c["b"]) if call > 0 else torch.ones(c["b"])
            grad = torch.autograd.grad(val, X, grad_outputs=go)[0]
            rnd = pr.input_transform["round"]
            rec = dict(X=X.detach().clone(), grad_output=go, value=val.detach(), grad=grad,
                       ma_pre=pre, ma_post=(float(pr.ma_counter), float(pr.ma_hidden)))
            # recover internals through the transforms (same objects the Function used)
            with torch.no_grad():
                tilde = pr.input_transform(X.detach().unsqueeze(-3))
                xun = pr.input_transform["unnormalize"](X.detach()) if "unnormalize" in pr.input_transform else X.detach()
                rec["tilde_x"] = tilde
                rec["prob"] = rnd.get_rounding_prob(xun)
                tx = pr.one_hot_to_numeric(tilde) if pr.one_hot_to_numeric is not None else tilde
                rec["acq_values"] = acq(tx)
            out["calls"].append(rec)
        out["base_samples"] = getattr(rnd, "base_samples", None)
        out["base_samples_categorical"] = getattr(rnd, "base_samples_categorical", None)
    # final discrete draw per restart (probabilistic_reparameterization.py:198-233) under a fixed seed of torch's global
    # (CPU) generator: the reference's own call sequence, one Bernoulli draw per integer parameter, one Categorical per block
    sc_seed = c["seed"] + 31
    torch.manual_seed(sc_seed)
    with torch.no_grad():
        sc_out = pr.sample_candidates(theta.clone())
    out["sample_candidates"] = dict(seed=sc_seed, X=theta.clone(), out=sc_out.clone())
    return out


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, c in make_cases().items():
        out = run_case(name, c)
        path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
        out.pop("spec", None)
        torch.save(out, path)
        print(f"{name}: wrote {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
