"""Helpers shared by the golden-vector tests: load a fixture and rebuild the oracle objects it describes."""
import os

import torch

from oracle import gp_oracle as G
from oracle import pr_oracle as P

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# named_*: the same shapes at the sizes the named configs run at (n_train = 119 / 219, mc = 128, 5 or 20 restarts)
MC_CASES = ["ackley13_ei", "rosenbrock10_ucb", "chemistry_mc_ei", "mixed_int_f1_ei", "neg_int_onehot_ei",
            "named_ackley13_ei", "named_rosenbrock10_ucb", "named_mixed_int_f1_ei", "chemistry_latent_mc_ei"]
ANALYTIC_CASES = ["chemistry_analytic_ei", "int_cat_analytic_pm", "named_chemistry_analytic_ei",
                  "chemistry_latent_analytic_ei"]


def load(name):
    return torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)


def oracle_space(g):
    c = g["case"]
    return P.PRSpace(dim=c["dim"], integer_indices=c["integer_indices"],
                     integer_bounds=torch.tensor(c["integer_bounds"], dtype=torch.float64).reshape(2, -1),
                     categorical_features=c["categorical_features"], tau=0.1)


def oracle_spec(g):
    c = g["case"]
    cf = c["categorical_features"] or {}
    dk = g["train_X_model"].shape[-1]
    if c["kernel_type"] == "mixed_latent":
        ctf = g["latent"]["ctf"]
    else:
        ctf = cf if c["kernel_type"] == "mixed_categorical" else {}
    return G.reference_kernel_spec(c["kernel_type"], dk, c["binary_dims"], ctf, **g["hp"])


def oracle_latent_transform(g):
    """mixed_latent cases: the oracle's restatement of the model-level embedding, with the fixture's tables."""
    lat = g.get("latent")
    if lat is None:
        return None
    return G.latent_embedding_transform(lat["tables"], [s[0] for s in lat["specs"]], g["train_X_numeric"].shape[-1])


def oracle_acq(g):
    c = g["case"]
    model = G.OracleGP(g["train_X_model"], g["train_Y"], oracle_spec(g), g["noise"], g["mean_const"],
                       input_transform=oracle_latent_transform(g))
    if c["acq"] == "ei":
        return G.OracleEI(model, best_f=g["train_Y"].max())
    if c["acq"] == "ucb":
        return G.OracleUCB(model, beta=c["beta"])
    return G.OraclePosteriorMean(model)
