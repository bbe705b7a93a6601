"""Build the PRODUCT-side objects (bo_pr_b200 classes on a CUDA device) that a golden fixture describes, so the GPU
parity tests read like the reference's own usage: get_kernel -> FixedNoiseGP -> ExpectedImprovement ->
MCProbabilisticReparameterization."""
import torch

import bo_pr_b200 as B


def _unique_leaves(k, out=None):
    out = [] if out is None else out
    name = type(k).__name__
    if name == "ScaleKernel":
        _unique_leaves(k.base_kernel, out)
    elif name in ("ProductKernel", "AdditiveKernel"):
        for m in k.kernels:
            _unique_leaves(m, out)
    elif not any(k is o for o in out):
        out.append(k)
    return out


def _scales(k, out=None):
    out = [] if out is None else out
    name = type(k).__name__
    if name == "ScaleKernel":
        out.append(k)
    elif name in ("ProductKernel", "AdditiveKernel"):
        for m in k.kernels:
            _scales(m, out)
    return out


def product_covar(case, hp, train_Xk, train_Y, ctf=None):
    cf = case["categorical_features"] or {}
    dk = train_Xk.shape[-1]
    covar = B.get_kernel(case["kernel_type"], dk, case["binary_dims"],
                         ctf if ctf is not None else (cf if case["kernel_type"] == "mixed_categorical" else {}),
                         train_Xk, train_Y)
    latent_starts = list(ctf.keys()) if ctf is not None else []
    for leaf in _unique_leaves(covar):
        dims = leaf.active_dims.tolist()
        if ctf is not None and dims[0] in latent_starts and len(dims) == ctf[dims[0]]:
            leaf.lengthscale = [hp["latent_lengthscale"][latent_starts.index(dims[0])]]
        elif type(leaf).__name__ == "CategoricalKernel":
            leaf.lengthscale = hp["categorical_lengthscale"]
        elif len(case["binary_dims"]) > 0 and dims == list(case["binary_dims"]):
            leaf.lengthscale = hp["binary_lengthscale"]
        else:
            leaf.lengthscale = hp["cont_lengthscale"]
    scales = _scales(covar)
    scales[0].outputscale = hp["outputscale"]
    if len(scales) > 1:
        scales[1].outputscale = hp["outputscale_sum"]
    return covar


def product_model(g, device):
    c = g["case"]
    Xk = g["train_X_model"].to(device)
    y = g["train_Y"].to(device)
    lat = g.get("latent")
    if lat is not None:
        # kernel_type "mixed_latent" as experiment_utils.py:289-313 builds it: numeric training inputs + a model-level
        # LatentCategoricalEmbedding whose (fitted) tables come from the fixture
        Xn = g["train_X_numeric"].to(device)
        emb = B.LatentCategoricalEmbedding([B.LatentCategoricalSpec(idx=i, num_categories=k, latent_dim=e)
                                            for i, k, e in lat["specs"]], dim=Xn.shape[-1]).to(device)
        with torch.no_grad():
            for (i, _, _), t in zip(lat["specs"], lat["tables"]):
                getattr(emb, f"raw_latent_emb_tables_{i}").copy_(t.to(device))
        covar = product_covar(c, g["hp"], Xk, y, ctf=lat["ctf"])
        return B.FixedNoiseGP(Xn, y.unsqueeze(-1), torch.full_like(y, g["noise"]).unsqueeze(-1), covar_module=covar,
                              input_transform=emb, mean_const=g["mean_const"])
    covar = product_covar(c, g["hp"], Xk, y)
    return B.FixedNoiseGP(Xk, y.unsqueeze(-1), torch.full_like(y, g["noise"]).unsqueeze(-1), covar_module=covar,
                          mean_const=g["mean_const"])


def product_acq(g, device, model=None):
    c = g["case"]
    model = product_model(g, device) if model is None else model
    if c["acq"] == "ei":
        return B.ExpectedImprovement(model, best_f=g["train_Y"].max())
    if c["acq"] == "ucb":
        return B.UpperConfidenceBound(model, beta=c["beta"])
    return B.PosteriorMean(model)


def integer_bounds(c, device):
    return torch.tensor(c["integer_bounds"], dtype=torch.float64, device=device).reshape(2, -1)


def product_mc_pr(g, device, acq=None):
    c = g["case"]
    acq = product_acq(g, device) if acq is None else acq
    pr = B.MCProbabilisticReparameterization(
        acq_function=acq, dim=c["dim"], integer_indices=c["integer_indices"], integer_bounds=integer_bounds(c, device),
        categorical_features=c["categorical_features"], batch_limit=16, apply_numeric=c["apply_numeric"],
        mc_samples=c["mc"], grad_estimator=c["grad_estimator"], tau=0.1)
    rnd = pr.input_transform["round"]
    # the fixture's base uniforms are the ones the reference drew (scrambled Sobol): install them as the cached buffers
    if g["base_samples"] is not None:
        rnd.register_buffer("base_samples", g["base_samples"].to(device).contiguous())
    if g["base_samples_categorical"] is not None:
        rnd.register_buffer("base_samples_categorical", g["base_samples_categorical"].to(device).contiguous())
    return pr


def product_analytic_pr(g, device, acq=None):
    c = g["case"]
    acq = product_acq(g, device) if acq is None else acq
    return B.AnalyticProbabilisticReparameterization(
        acq_function=acq, dim=c["dim"], dtype=torch.float64, device=torch.device(device),
        integer_indices=c["integer_indices"], integer_bounds=integer_bounds(c, device),
        categorical_features=c["categorical_features"], apply_numeric=c["apply_numeric"], tau=0.1)


def close(a, b, rtol=1e-9, atol=1e-12):
    """|a - b| <= rtol |b| + atol  (the north-star tolerance: 1e-9 relative in float64, SURVEY.md section 7 'hard parts'
    for why an absolute floor is needed: sigma^2 and EI cancel catastrophically near training points)."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    err = (a - b).abs()
    ok = err <= rtol * b.abs() + atol
    return bool(ok.all()), float((err / (rtol * b.abs() + atol)).max()) if err.numel() else 0.0
