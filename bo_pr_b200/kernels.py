"""Mixed categorical / ordinal / continuous kernel: structure holders + the reference's factory.

Mirrors ``discrete_mixed_bo/kernels.py:18-101`` (``get_kernel``) and the slice of the GPyTorch / BoTorch kernel
surface it uses (``MaternKernel``, ``ScaleKernel``, ``ProductKernel``, ``AdditiveKernel``, ``CategoricalKernel``;
``*`` and ``+`` composition).  These classes only *describe* the covariance -- hyperparameters live in plain tensors
with the GPyTorch attribute names (``lengthscale`` ``[1, k]``, ``outputscale``, ``active_dims``, ``base_kernel``,
``kernels``) -- the arithmetic runs in ``libprb200`` (``kstar_panel_kernel`` / ``grad_contract_kernel``).
``kernel_terms`` flattens a kernel tree (these classes *or* duck-typed real GPyTorch kernels) into the additive
terms x multiplicative factors layout of ``prb_model_desc``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor


class Interval:
    """Stand-in for ``gpytorch.constraints.Interval`` (bounds are recorded and enforced on assignment)."""

    def __init__(self, lower_bound: float, upper_bound: float):
        self.lower_bound, self.upper_bound = float(lower_bound), float(upper_bound)


class Kernel:
    """Base class: ``*`` builds a flattened ProductKernel, ``+`` a flattened AdditiveKernel (GPyTorch semantics;
    the reference relies on it at ``kernels.py:85-94``)."""

    def __mul__(self, other: "Kernel") -> "ProductKernel":
        ks = list(self.kernels) if isinstance(self, ProductKernel) else [self]
        ks += list(other.kernels) if isinstance(other, ProductKernel) else [other]
        return ProductKernel(*ks)

    def __add__(self, other: "Kernel") -> "AdditiveKernel":
        ks = list(self.kernels) if isinstance(self, AdditiveKernel) else [self]
        ks += list(other.kernels) if isinstance(other, AdditiveKernel) else [other]
        return AdditiveKernel(*ks)


class _StationaryKernel(Kernel):
    def __init__(self, ard_num_dims: Optional[int] = None, active_dims: Optional[Sequence[int]] = None,
                 lengthscale_constraint: Optional[Interval] = None, lengthscale_prior=None, **kwargs):
        self.ard_num_dims = ard_num_dims
        self.active_dims = None if active_dims is None else torch.as_tensor(list(active_dims), dtype=torch.long)
        self.lengthscale_constraint = lengthscale_constraint
        k = 1 if ard_num_dims is None else int(ard_num_dims)
        self._lengthscale = torch.full((1, k), 0.6931471805599453, dtype=torch.float64)  # softplus(0), GPyTorch's init

    @property
    def lengthscale(self) -> Tensor:
        return self._lengthscale

    @lengthscale.setter
    def lengthscale(self, value) -> None:
        v = torch.as_tensor(value, dtype=torch.float64).reshape(1, -1)
        if v.shape[-1] != self._lengthscale.shape[-1]:
            raise ValueError(f"lengthscale has {v.shape[-1]} entries, kernel expects {self._lengthscale.shape[-1]}")
        c = self.lengthscale_constraint
        if c is not None and bool(((v < c.lower_bound) | (v > c.upper_bound)).any()):
            raise ValueError(f"lengthscale {v.tolist()} violates Interval({c.lower_bound}, {c.upper_bound})")
        self._lengthscale = v.clone()
        self._version = getattr(self, "_version", 0) + 1


class MaternKernel(_StationaryKernel):
    """``gpytorch.kernels.MaternKernel``; only ``nu=2.5`` exists on the device."""

    def __init__(self, nu: float = 2.5, **kwargs):
        if nu != 2.5:
            raise NotImplementedError("libprb200 implements the Matern-5/2 kernel only (the reference uses nu=2.5).")
        super().__init__(**kwargs)
        self.nu = nu


class CategoricalKernel(_StationaryKernel):
    """``botorch.models.kernels.CategoricalKernel``: exp(-mean_j([x_j != x'_j] / lengthscale_j))."""


class ScaleKernel(Kernel):
    def __init__(self, base_kernel: Kernel, outputscale_prior=None, outputscale_constraint=None, **kwargs):
        self.base_kernel = base_kernel
        self._outputscale = torch.tensor(0.6931471805599453, dtype=torch.float64)

    @property
    def outputscale(self) -> Tensor:
        return self._outputscale

    @outputscale.setter
    def outputscale(self, value) -> None:
        v = torch.as_tensor(value, dtype=torch.float64).reshape(())
        if not bool(v > 0):
            raise ValueError("outputscale must be positive")
        self._outputscale = v.clone()
        self._version = getattr(self, "_version", 0) + 1


class ProductKernel(Kernel):
    def __init__(self, *kernels: Kernel):
        self.kernels = list(kernels)


class AdditiveKernel(Kernel):
    def __init__(self, *kernels: Kernel):
        self.kernels = list(kernels)


# ----------------------------------------------------------------------------------------------------------------------
def get_kernel(
    kernel_type: str,
    dim: int,
    binary_dims: List[int],
    categorical_transformed_features: Dict[int, int],
    train_X: Tensor,
    train_Y: Tensor,
    function_name: Optional[str] = None,
    use_ard_binary: bool = False,
) -> Optional[Kernel]:
    """Same signature and structure as the reference's ``get_kernel`` (``kernels.py:18-101``).

    ``"mixed"``: ``Scale(M52_ARD(non-binary dims) * M52(binary dims))``; ``"mixed_categorical"``:
    ``Scale(prod) + Scale(sum)`` where -- replicating the reference's second ``prod_kernel *= k`` loop
    (``kernels.py:90-94``) -- every factor but the first enters the product squared.  ``"mixed_latent"``
    (``kernels.py:73-84``): ``Scale(M52_ARD(continuous) * M52(binary) * prod_j M52_iso(latent columns of categorical j))``
    over the ``LatentCategoricalEmbedding`` layout, ``categorical_transformed_features = {first latent column: latent
    dim}``.  ``"ard_combo"`` and ``"iso_combo"`` are outside the hot-path scope and raise.
    """
    if kernel_type == "mixed_categorical":
        categorical_dims = list(categorical_transformed_features.keys())
    else:
        if len(categorical_transformed_features) > 0:
            start = min(categorical_transformed_features.keys())
            categorical_dims = list(range(start, dim))
        else:
            categorical_dims = []
    if kernel_type in ("mixed", "mixed_categorical", "mixed_latent"):
        cont_dims = sorted(set(range(dim)) - set(binary_dims))
        if kernel_type in ("mixed_categorical", "mixed_latent"):
            cont_dims = sorted(set(cont_dims) - set(categorical_dims))
        kernels: List[Kernel] = []
        if len(cont_dims) > 0:
            kernels.append(MaternKernel(nu=2.5, ard_num_dims=len(cont_dims), active_dims=cont_dims,
                                        lengthscale_constraint=Interval(0.1, 20.0)))
        if len(binary_dims) > 0:
            kernels.append(MaternKernel(nu=2.5, ard_num_dims=len(binary_dims) if use_ard_binary else None,
                                        active_dims=binary_dims, lengthscale_constraint=Interval(0.1, 20.0)))
        if kernel_type == "mixed_categorical" and len(categorical_dims) > 0:
            kernels.append(CategoricalKernel(ard_num_dims=len(categorical_dims), active_dims=categorical_dims,
                                             lengthscale_constraint=Interval(1e-3, 20.0)))
        if kernel_type == "mixed_latent":
            for start, latent_dim in categorical_transformed_features.items():
                # one isotropic kernel per set of latent embedding columns
                kernels.append(MaternKernel(nu=2.5, ard_num_dims=None,
                                            active_dims=list(range(start, start + latent_dim)),
                                            lengthscale_constraint=Interval(1e-3, 20.0)))
        prod_kernel = kernels[0]
        for k in kernels[1:]:
            prod_kernel = prod_kernel * k
        if kernel_type != "mixed_categorical":
            return ScaleKernel(prod_kernel)
        sum_kernel = kernels[0]
        for k in kernels[1:]:
            prod_kernel = prod_kernel * k  # the reference's double multiply: factors[1:] end up squared
            sum_kernel = sum_kernel + k
        return ScaleKernel(prod_kernel) + ScaleKernel(sum_kernel)
    elif kernel_type == "botorch_default":
        return None
    elif kernel_type in ("ard_combo", "iso_combo"):
        raise NotImplementedError(
            f"kernel_type={kernel_type!r} is outside the B200 hot-path scope (no named config uses it).")
    raise ValueError(f"{kernel_type} is not supported.")


# ----------------------------------------------------------------------------------------------------------------------
# Kernel tree -> (terms x factors).  Works on the classes above and on duck-typed GPyTorch/BoTorch kernels.
# ----------------------------------------------------------------------------------------------------------------------
Factor = Tuple[str, List[int], List[float], int]  # (kind, dims, lengthscales, power)
TermT = Tuple[float, bool, List[Factor]]  # (outputscale, additive, factors)


def _kind(k) -> Optional[str]:
    name = type(k).__name__
    if name == "MaternKernel":
        if float(getattr(k, "nu", 2.5)) != 2.5:
            raise NotImplementedError("only Matern nu=2.5 is implemented on the device")
        return "matern52"
    if name == "CategoricalKernel":
        return "categorical"
    return None


def _leaf(k, d_k: int) -> Tuple[str, List[int], List[float]]:
    kind = _kind(k)
    if kind is None:
        raise NotImplementedError(f"kernel {type(k).__name__} has no device implementation in libprb200")
    dims = list(range(d_k)) if getattr(k, "active_dims", None) is None else [int(i) for i in k.active_dims]
    ls = torch.as_tensor(k.lengthscale).detach().reshape(-1).to(torch.float64).cpu().tolist()
    if len(ls) == 1:
        ls = ls * len(dims)
    if len(ls) != len(dims):
        raise ValueError("lengthscale / active_dims mismatch")
    return kind, dims, ls


def _factors(k, d_k: int, additive: bool) -> List[Factor]:
    name = type(k).__name__
    members = None
    if name == "ProductKernel" and not additive:
        members = list(k.kernels)
    elif name == "AdditiveKernel" and additive:
        members = list(k.kernels)
    if members is None:
        members = [k]
    out: List[Factor] = []
    seen: List[object] = []
    for m in members:
        if any(m is s for s in seen):  # the same object listed again => one more power (kernels.py:90-94)
            idx = [i for i, s in enumerate(seen) if s is m][0]
            kind, dims, ls, p = out[idx]
            if additive:
                raise NotImplementedError("repeated member of an AdditiveKernel")
            out[idx] = (kind, dims, ls, p + 1)
            continue
        kind, dims, ls = _leaf(m, d_k)
        out.append((kind, dims, ls, 1))
        seen.append(m)
    return out


def kernel_terms(covar_module, d_k: int) -> List[TermT]:
    """Flatten ``Scale(Product|leaf)``, ``Scale(Additive)``, sums of those, or a bare leaf/product."""
    name = type(covar_module).__name__
    tops = list(covar_module.kernels) if name == "AdditiveKernel" and all(
        type(k).__name__ == "ScaleKernel" for k in covar_module.kernels) else [covar_module]
    terms: List[TermT] = []
    for top in tops:
        if type(top).__name__ == "ScaleKernel":
            scale = float(torch.as_tensor(top.outputscale).detach().reshape(-1)[0])
            base = top.base_kernel
        else:
            scale, base = 1.0, top
        additive = type(base).__name__ == "AdditiveKernel"
        terms.append((scale, additive, _factors(base, d_k, additive)))
    return terms


def kernel_fingerprint(k) -> tuple:
    """Cheap identity of a kernel tree's hyper-parameters: the classes above bump ``_version`` on every assignment;
    duck-typed GPyTorch kernels are fingerprinted through the storage and in-place version counter of their (raw)
    parameter tensors.  ``GPState`` caches are rebuilt when this changes (stale Cholesky factors would silently give wrong
    PR values)."""
    name = type(k).__name__
    out = [name, id(k), getattr(k, "_version", 0)]
    for attr in ("raw_lengthscale", "raw_outputscale"):
        t = getattr(k, attr, None)
        if torch.is_tensor(t):
            out += [t.data_ptr(), t._version]
    if name == "ScaleKernel":
        out.append(kernel_fingerprint(k.base_kernel))
    elif name in ("ProductKernel", "AdditiveKernel"):
        out += [kernel_fingerprint(m) for m in k.kernels]
    return tuple(out)
