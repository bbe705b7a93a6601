"""Seeded random PR problems in the reference's column-order contract (continuous | integers | one-hot blocks) for the
differential tests: search space, kernel structure as ``get_kernel`` would build it, GP data and hyper-parameters."""
import math

import torch

from oracle import gp_oracle as G


def random_case(seed: int):
    g = torch.Generator().manual_seed(seed)

    def ri(lo, hi):
        return int(torch.randint(lo, hi + 1, (1,), generator=g))

    n_cont = ri(0, 4)
    n_int = ri(0, 5)
    n_cat = ri(0, 3) if seed % 3 != 0 else 0
    if n_int + n_cat == 0:
        n_int = 2
    cards = [ri(2, 5) for _ in range(n_cat)]
    int_lo, int_hi = [], []
    for _ in range(n_int):
        kind = ri(0, 3)
        if kind == 0:      # binary
            lo, hi = 0, 1
        elif kind == 1:    # ordinal from 0
            lo, hi = 0, ri(2, 6)
        elif kind == 2:    # negative lower bound
            lo = -ri(1, 3)
            hi = lo + ri(1, 5)
        else:
            lo = ri(1, 3)
            hi = lo + ri(1, 4)
        int_lo.append(float(lo))
        int_hi.append(float(hi))
    int_idx = list(range(n_cont, n_cont + n_int))
    start = n_cont + n_int
    dim = start + sum(cards)
    use_numeric = n_cat > 0 and ri(0, 1) == 1
    kernel_type = "mixed_categorical" if use_numeric else "mixed"
    cf = {start + j: c for j, c in enumerate(cards)} if n_cat > 0 else None
    binary_dims = [c for c, lo, hi in zip(int_idx, int_lo, int_hi) if hi - lo == 1.0]
    if kernel_type == "mixed" and len(binary_dims) == len(range(dim)):
        binary_dims = binary_dims[:-1]
    case = dict(dim=dim, integer_indices=int_idx, integer_bounds=[int_lo, int_hi], categorical_features=cf,
                kernel_type=kernel_type, binary_dims=binary_dims, n_train=ri(3, 70), b=ri(1, 6), mc=ri(1, 70),
                acq=["ei", "ucb", "pm"][ri(0, 2)], beta=0.2 * dim * math.log(2 * ri(1, 20)), apply_numeric=use_numeric,
                grad_estimator=["reinforce", "reinforce_ma"][ri(0, 1)], seed=seed)
    return case


def build_problem(c):
    """GP data + hyper-parameters for a case (same construction as oracle/gen_golden.py::build_problem)."""
    g = torch.Generator().manual_seed(c["seed"] + 77)
    dim, cf, ints = c["dim"], c["categorical_features"] or {}, c["integer_indices"]
    ib = torch.tensor(c["integer_bounds"], dtype=torch.float64).reshape(2, -1)
    n = c["n_train"]
    X = torch.rand(n, dim, generator=g, dtype=torch.float64)
    for k, col in enumerate(ints):
        lo, hi = ib[0, k], ib[1, k]
        v = torch.floor(lo + X[:, col] * (hi - lo + 1)).clamp(lo, hi)
        X[:, col] = (v - lo) / (hi - lo)
    if cf:
        s = min(cf.keys())
        for _, card in cf.items():
            cat = torch.randint(0, card, (n,), generator=g)
            X[:, s:s + card] = torch.nn.functional.one_hot(cat, card).to(X)
            s += card
    if c["apply_numeric"]:
        start = min(cf.keys())
        cols, s = [X[:, :start]], start
        for _, card in cf.items():
            cols.append(X[:, s:s + card].argmax(-1, keepdim=True).to(X))
            s += card
        Xk = torch.cat(cols, dim=-1)
    else:
        Xk = X
    dk = Xk.shape[-1]
    binary = c["binary_dims"]
    cat_dims = list(cf.keys()) if c["kernel_type"] == "mixed_categorical" else []
    cont_dims = sorted(set(range(dk)) - set(binary) - set(cat_dims))
    hp = dict(cont_lengthscale=(0.3 + 0.9 * torch.rand(len(cont_dims), generator=g, dtype=torch.float64)).tolist(),
              binary_lengthscale=[0.5 + float(torch.rand(1, generator=g))] if binary else [],
              categorical_lengthscale=(0.5 + 1.5 * torch.rand(len(cat_dims), generator=g, dtype=torch.float64)).tolist(),
              outputscale=0.5 + float(torch.rand(1, generator=g)), outputscale_sum=0.2 + float(torch.rand(1, generator=g)))
    spec = G.reference_kernel_spec(c["kernel_type"], dk, binary, cf if c["kernel_type"] == "mixed_categorical" else {}, **hp)
    K = G.eval_kernel(spec, Xk, Xk) + 1e-6 * torch.eye(n, dtype=torch.float64)
    y = torch.linalg.cholesky(K) @ torch.randn(n, generator=g, dtype=torch.float64)
    if n > 1:
        y = (y - y.mean()) / y.std()
    theta = torch.rand(c["b"], 1, dim, generator=g, dtype=torch.float64)
    if c["b"] > 1 and len(ints) > 0:  # edge rows: exact bounds
        theta[0, 0, ints] = 1.0
        theta[1, 0, ints] = 0.0
    mc = c["mc"]
    from torch.quasirandom import SobolEngine
    bs = SobolEngine(len(ints), scramble=True, seed=c["seed"]).draw(mc, dtype=torch.float64).view(mc, 1, -1) if ints else None
    bc = SobolEngine(len(cf), scramble=True, seed=c["seed"] + 1).draw(mc, dtype=torch.float64).view(mc, 1, -1) if cf else None
    return dict(case=c, theta=theta, train_X_model=Xk, train_Y=y, hp=hp, noise=1e-6,
                mean_const=float(torch.randn(1, generator=g)) * 0.3, cont_dims=cont_dims, cat_dims=cat_dims,
                base_samples=bs, base_samples_categorical=bc)
