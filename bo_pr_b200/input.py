"""PR input transforms backed by the CUDA sampler.

API mirror of ``discrete_mixed_bo/input.py`` for the PR path: ``OneHotToNumeric`` (30-117), ``LatentCategoricalEmbedding``
(766-952; inside a PR call the device sampler performs the table lookup, see ``prb_space_desc.latent_tables``),
``AnalyticProbabilisticReparameterizationInputTransform`` (245-502), ``MCProbabilisticReparameterizationInputTransform``
(505-763), plus the BoTorch ``InputTransform`` / ``Normalize`` / ``ChainedInputTransform`` surface they plug into.
The discrete sampling, probabilities and enumeration run in ``libprb200`` (``pr_sample_kernel``,
``analytic_build_kernel``); ``Normalize`` and ``OneHotToNumeric`` applied on their own are trivial torch element-wise
ops on the device (inside the fused PR call the kernels do both).
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor
from torch.nn import Module, ModuleDict
from torch.quasirandom import SobolEngine

from . import _capi as capi


# ----------------------------------------------------------------------------------------------------------------------
# BoTorch transform surface
# ----------------------------------------------------------------------------------------------------------------------
class InputTransform:
    transform_on_train: bool = True
    transform_on_eval: bool = True
    transform_on_fantasize: bool = True

    def forward(self, X: Tensor) -> Tensor:
        if self.training:
            if self.transform_on_train:
                return self.transform(X)
        elif self.transform_on_eval:
            return self.transform(X)
        return X

    def transform(self, X: Tensor) -> Tensor:  # pragma: no cover - abstract
        raise NotImplementedError

    def equals(self, other) -> bool:
        return type(self) is type(other)


class Normalize(InputTransform, Module):
    """``botorch.models.transforms.input.Normalize`` restricted to what the reference uses
    (``probabilistic_reparameterization.py:54-63, 84-93``): explicit ``bounds``, optional ``indices``, ``reverse``."""

    def __init__(self, d: int, indices=None, bounds: Optional[Tensor] = None, transform_on_train: bool = True,
                 transform_on_eval: bool = True, transform_on_fantasize: bool = True, reverse: bool = False, **kwargs):
        super().__init__()
        if bounds is None:
            raise NotImplementedError("learned bounds are not supported")
        if indices is not None and len(indices) > 0:
            self.register_buffer("indices", torch.as_tensor(indices, dtype=torch.long, device=bounds.device))
        self.register_buffer("mins", bounds[0:1, :].clone())
        self.register_buffer("ranges", (bounds[1:2, :] - bounds[0:1, :]).clone())
        self.transform_on_train = transform_on_train
        self.transform_on_eval = transform_on_eval
        self.transform_on_fantasize = transform_on_fantasize
        self.reverse = reverse

    def _cols(self, X: Tensor, fn) -> Tensor:
        if hasattr(self, "indices"):
            X_new = X.clone()
            X_new[..., self.indices] = fn(X[..., self.indices], self.mins[..., self.indices],
                                          self.ranges[..., self.indices])
            return X_new
        return fn(X, self.mins, self.ranges)

    def transform(self, X: Tensor) -> Tensor:
        if self.reverse:
            return self._cols(X, lambda x, m, r: m + x * r)
        return self._cols(X, lambda x, m, r: (x - m) / r)

    def untransform(self, X: Tensor) -> Tensor:
        if self.reverse:
            return self._cols(X, lambda x, m, r: (x - m) / r)
        return self._cols(X, lambda x, m, r: m + x * r)


class ChainedInputTransform(InputTransform, ModuleDict):
    def __init__(self, **transforms):
        super().__init__(OrderedDict(transforms))
        self.transform_on_train = any(tf.transform_on_train for tf in transforms.values())
        self.transform_on_eval = any(tf.transform_on_eval for tf in transforms.values())
        self.transform_on_fantasize = any(tf.transform_on_fantasize for tf in transforms.values())

    def transform(self, X: Tensor) -> Tensor:
        for tf in self.values():
            X = tf.forward(X)
        return X


# ----------------------------------------------------------------------------------------------------------------------
def _categorical_layout(categorical_features: Optional[Dict[int, int]]):
    """starts / ends of the one-hot blocks (``input.py:583-594``): the first key is the first one-hot column, blocks
    are contiguous in dict order."""
    starts, ends = [], []
    start = None
    for i, card in (categorical_features or {}).items():
        if start is None:
            start = i
        starts.append(start)
        ends.append(start + card)
        start = start + card
    return starts, ends


def make_space_desc(dim: int, integer_indices, integer_bounds: Optional[Tensor],
                    categorical_features: Optional[Dict[int, int]], tau: float, apply_numeric: bool = False,
                    raw_integer_space: bool = False, latent: Optional["LatentCategoricalEmbedding"] = None,
                    device=None) -> capi.PrbSpaceDesc:
    ints = [int(i) for i in (integer_indices if integer_indices is not None else [])]
    starts, ends = _categorical_layout(categorical_features)
    if dim > capi.PRB_MAX_DIMS:
        raise NotImplementedError(f"dim {dim} > {capi.PRB_MAX_DIMS} is not supported by libprb200")
    if len(starts) > capi.PRB_MAX_CAT:
        raise NotImplementedError(f"more than {capi.PRB_MAX_CAT} categorical parameters")
    sp = capi.PrbSpaceDesc()
    sp.d, sp.n_int, sp.n_cat = int(dim), len(ints), len(starts)
    sp.apply_numeric = int(bool(apply_numeric) and len(starts) > 0)
    sp.tau = float(tau)
    sp.raw_integer_space = int(bool(raw_integer_space))
    if len(ints) > 0:
        ib = torch.as_tensor(integer_bounds).detach().to("cpu", torch.float64).reshape(2, -1)
        for k, c in enumerate(ints):
            sp.int_cols[k], sp.int_lo[k], sp.int_hi[k] = c, float(ib[0, k]), float(ib[1, k])
    for j, (s, e) in enumerate(zip(starts, ends)):
        if e - s > capi.PRB_MAX_CARD:
            raise NotImplementedError(f"categorical cardinality {e - s} > {capi.PRB_MAX_CARD}")
        sp.cat_start[j], sp.cat_card[j] = s, e - s
    if latent is not None:
        if not sp.apply_numeric:
            raise NotImplementedError("a latent categorical embedding needs apply_numeric=True (OneHotToNumeric first)")
        want = [starts[0] + j for j in range(len(starts))]
        if latent.categ_idcs.tolist() != want or [c.num_categories for c in latent._specs] != [e - s for s, e in
                                                                                                zip(starts, ends)]:
            raise ValueError("the latent embedding's categorical specs do not match the numeric layout of the search space")
        flat, dims = latent.device_tables(device)
        for j, e in enumerate(dims):
            sp.latent_dim[j] = int(e)
        sp.latent_tables = flat.data_ptr()
        sp._keepalive = flat  # the descriptor points into this tensor
    return sp


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class OneHotToNumeric(InputTransform, Module):
    """``input.py:30-117``: one-hot blocks (trailing columns) -> one numeric code column each."""

    def __init__(self, categorical_features: Optional[Dict[int, int]] = None, transform_on_train: bool = False,
                 transform_on_eval: bool = True, transform_on_fantasize: bool = False, use_ste: bool = False):
        super().__init__()
        if use_ste:
            raise NotImplementedError("straight-through estimators are a competing baseline, not part of the PR path")
        self.transform_on_train = transform_on_train
        self.transform_on_eval = transform_on_eval
        self.transform_on_fantasize = transform_on_fantasize
        self.categorical_features = (None if (categorical_features is None or len(categorical_features) == 0)
                                     else categorical_features)
        self.categorical_starts: List[int] = []
        self.categorical_ends: List[int] = []
        if self.categorical_features is not None:
            start_idx = None
            for i in sorted(categorical_features.keys()):
                if start_idx is None:
                    start_idx = i
                self.categorical_starts.append(start_idx)
                end_idx = start_idx + categorical_features[i]
                self.categorical_ends.append(end_idx)
                start_idx = end_idx
            self.numeric_dim = min(self.categorical_starts) + len(categorical_features)

    def transform(self, X: Tensor) -> Tensor:
        if self.categorical_features is None:
            return X
        X_numeric = X[..., : self.numeric_dim].clone()
        idx = self.categorical_starts[0]
        for start, end in zip(self.categorical_starts, self.categorical_ends):
            X_numeric[..., idx] = X[..., start:end].argmax(dim=-1).to(X)
            idx += 1
        return X_numeric

    def untransform(self, X: Tensor) -> Tensor:
        if X.requires_grad:
            raise NotImplementedError
        if self.categorical_features is None:
            return X
        one_hots = [torch.nn.functional.one_hot(X[..., idx].long(), num_classes=card).to(X)
                    for idx, card in sorted(self.categorical_features.items(), key=lambda kv: kv[0])]
        return torch.cat([X[..., : min(self.categorical_features.keys())], *one_hots], dim=-1)


# ----------------------------------------------------------------------------------------------------------------------
# Latent categorical embedding (kernel_type "mixed_latent")
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class CategoricalSpec:
    """``input.py:766-769``."""
    idx: int
    num_categories: int


@dataclass
class LatentCategoricalSpec(CategoricalSpec):
    """``input.py:772-774``."""
    latent_dim: int


class LatentCategoricalEmbedding(InputTransform, Module):
    """``input.py:777-952``: every categorical column (an integer code, i.e. the ``OneHotToNumeric`` layout) is replaced by
    the row of a learned ``num_categories x latent_dim`` table; the output is ``[non-categorical columns, embeddings in
    spec order]``.  The tables are parameters named ``raw_latent_emb_tables_{idx}`` as in the reference (fitted with the
    GP hyper-parameters, outside the hot path); ``get_emb_table`` pins row 0 to the origin and entry ``[1, 0]`` to zero
    (``input.py:902-909``).  Applied on its own this is a torch gather; inside a PR call the sampler kernel writes the
    embedding rows straight into the kernel input (``device_tables`` feeds ``prb_space_desc.latent_tables``)."""

    def __init__(self, categorical_specs: Sequence[LatentCategoricalSpec], dim: int, transform_on_train: bool = True,
                 transform_on_eval: bool = True, transform_on_preprocess: bool = False,
                 transform_on_fantasize: bool = False, eps: float = 1e-7) -> None:
        super().__init__()
        self._eps = eps
        self.dim = int(dim)
        self.transform_on_train = transform_on_train
        self.transform_on_eval = transform_on_eval
        self.transform_on_preprocess = transform_on_preprocess
        self.transform_on_fantasize = transform_on_fantasize
        self._emb_dim = 0
        self._specs: List[LatentCategoricalSpec] = []
        categ_idcs: List[int] = []
        for c in categorical_specs:
            idx = int(c.idx) + (self.dim if c.idx < 0 else 0)
            if not 0 <= idx < self.dim:
                raise ValueError(f"categorical index {c.idx} is out of range for dim {dim}")
            categ_idcs.append(idx)
            self._specs.append(LatentCategoricalSpec(idx=idx, num_categories=int(c.num_categories),
                                                     latent_dim=int(c.latent_dim)))
            # initial table: quasi-random standard-normal rows (the reference draws them with BoTorch's
            # draw_sobol_normal_samples; the values are a starting point for the fit, not part of the path's arithmetic)
            u = SobolEngine(int(c.latent_dim), scramble=True, seed=idx).draw(int(c.num_categories), dtype=torch.float64)
            v = 0.5 + (1.0 - 1e-10) * (u - 0.5)
            self.register_parameter(f"raw_latent_emb_tables_{idx}",
                                    torch.nn.Parameter(math.sqrt(2.0) * torch.erfinv(2.0 * v - 1.0)))
            self._emb_dim += int(c.latent_dim)
        self.register_buffer("categ_idcs", torch.tensor(categ_idcs, dtype=torch.long))
        mask = torch.ones(self.dim, dtype=torch.bool)
        mask[self.categ_idcs] = False
        self.register_buffer("non_categ_mask", mask)
        self._dev_tables: Optional[Tuple[tuple, Tensor]] = None

    def get_emb_table(self, idx: int) -> Tensor:
        raw = getattr(self, f"raw_latent_emb_tables_{idx}")
        with torch.no_grad():
            if bool((raw[0] != 0).any()):
                raw[0] = 0
            if raw.shape[1] > 1 and raw.shape[0] > 1 and bool(raw[1, 0] != 0):
                raw[1, 0] = 0
        return raw

    def transform(self, X: Tensor) -> Tensor:
        parts = [X[..., self.non_categ_mask]]
        for idx in self.categ_idcs.tolist():
            table = self.get_emb_table(idx).to(X)
            codes = X[..., idx].reshape(-1).long()
            parts.append(table.index_select(0, codes).reshape(*X.shape[:-1], table.shape[-1]))
        return torch.cat(parts, dim=-1)

    def transform_bounds(self, bounds: Tensor) -> Tensor:
        d_cont = self.dim - self.categ_idcs.shape[0]
        out = torch.zeros(2, d_cont + self._emb_dim, dtype=bounds.dtype, device=bounds.device)
        out[:, :d_cont] = bounds[:, self.non_categ_mask]
        out[1, d_cont:] = 1
        return out

    def untransform(self, X: Tensor, dist_func: Optional[Callable[[Tensor, Tensor, int], Tensor]] = None) -> Tensor:
        """Nearest-embedding decoding back to integer codes (not differentiable), ``input.py:911-952``."""
        out = torch.empty(*X.shape[:-1], self.dim, dtype=X.dtype, device=X.device)
        n_plain = X.shape[-1] - self._emb_dim
        out[..., self.non_categ_mask] = X[..., :n_plain]
        start = self.dim - self.categ_idcs.shape[0]
        for idx in self.categ_idcs.tolist():
            table = self.get_emb_table(idx).to(X)
            end = start + table.shape[-1]
            x = X[..., start:end].unsqueeze(-2)
            t = table.unsqueeze(-3)
            dist = dist_func(x, t, start) if dist_func is not None else torch.norm(x - t, dim=-1)
            out[..., idx] = dist.argmin(dim=-1).to(X.dtype)
            start = end
        return out

    # ---- device side ------------------------------------------------------------------------------------------------
    def fingerprint(self) -> tuple:
        out = []
        for idx in self.categ_idcs.tolist():
            t = getattr(self, f"raw_latent_emb_tables_{idx}")
            out += [t.data_ptr(), t._version]
        return tuple(out)

    def device_tables(self, device) -> Tuple[Tensor, List[int]]:
        """The pinned tables back to back (categorical order, row-major, float64) on ``device`` + their latent dims; cached
        until a table changes."""
        fp = (str(device),) + self.fingerprint()
        if self._dev_tables is None or self._dev_tables[0] != fp:
            flat = torch.cat([self.get_emb_table(i).detach().to(device=device, dtype=torch.float64).reshape(-1)
                              for i in self.categ_idcs.tolist()]).contiguous()
            fp = (str(device),) + self.fingerprint()  # get_emb_table may have pinned entries in place
            self._dev_tables = (fp, flat)
        return self._dev_tables[1], [s.latent_dim for s in self._specs]


# ----------------------------------------------------------------------------------------------------------------------
class _PRTransformBase(InputTransform, Module):
    def _setup(self, integer_indices, integer_bounds, categorical_features, tau):
        if integer_indices is None and categorical_features is None:
            raise ValueError("integer_indices and/or categorical_features must be provided.")
        discrete_indices: List[int] = []
        if integer_indices is not None and len(integer_indices) > 0:
            assert integer_bounds is not None
            self.register_buffer("integer_indices", torch.tensor(list(integer_indices), dtype=torch.long))
            self.register_buffer("integer_bounds", integer_bounds.clone())
            discrete_indices += list(integer_indices)
        else:
            self.integer_indices = None
            self.register_buffer("integer_bounds", torch.tensor([], dtype=torch.long))
        self.categorical_features = categorical_features
        starts, ends = _categorical_layout(categorical_features)
        for s, e in zip(starts, ends):
            discrete_indices += list(range(s, e))
        self.register_buffer("discrete_indices", torch.tensor(discrete_indices, dtype=torch.long))
        self.register_buffer("categorical_starts", torch.tensor(starts, dtype=torch.long))
        self.register_buffer("categorical_ends", torch.tensor(ends, dtype=torch.long))
        self.tau = tau
        self._lib = None

    @property
    def lib(self):
        if self._lib is None:
            self._lib = capi.load_library()
        return self._lib

    def _ints(self) -> List[int]:
        return [] if self.integer_indices is None else self.integer_indices.tolist()

    def _raw_space(self, dim: int, apply_numeric: bool = False) -> capi.PrbSpaceDesc:
        return make_space_desc(dim, self._ints(), self.integer_bounds if self.integer_indices is not None else None,
                               self.categorical_features, self.tau, apply_numeric=apply_numeric,
                               raw_integer_space=True)

    def get_rounding_prob(self, X: Tensor) -> Tensor:
        """``input.py:613-635`` on UN-normalised ``X [..., d]`` -> ``[..., n_disc]`` (integers first, then one-hots)."""
        if self.tau is None:
            raise NotImplementedError("tau=None is not used by the reference's PR path")
        capi.require_cuda(X, "X")
        lead, d = X.shape[:-1], X.shape[-1]
        th = X.detach().to(torch.float64).reshape(-1, d).contiguous()
        n_disc = int(self.discrete_indices.shape[0])
        prob = torch.empty(th.shape[0], n_disc, dtype=torch.float64, device=X.device)
        sp = self._raw_space(d)
        with torch.cuda.device(X.device):
            capi.check(self.lib.prb_pr_sample(C.byref(sp), th.data_ptr(), th.shape[0], 1, None, None, 0, 0, None,
                                              None, prob.data_ptr(), _stream(X.device)), "prb_pr_sample(prob)")
        # mc = 1 with Philox draws: only the probabilities are read back
        return prob.reshape(*lead, n_disc)


class MCProbabilisticReparameterizationInputTransform(_PRTransformBase):
    """``input.py:505-763``.  ``transform`` maps UN-normalised ``X [b, 1, q=1, d]`` to ``[b, mc, 1, d]``.

    Base uniforms (``base_samples [mc, 1, n_int]``, ``base_samples_categorical [mc, 1, n_cat]``) are created lazily,
    cached as buffers and shared by all batch elements, exactly like the reference (:655-673, 696-713): scrambled Sobol
    drawn on the CPU with a seed taken from torch's global generator -- or, with ``rng="philox"``, the library's
    counter-based Philox stream ``(philox_seed, philox_offset)`` generated on the device."""

    def __init__(self, integer_indices: Optional[List[int]] = None, integer_bounds: Optional[Tensor] = None,
                 categorical_features: Optional[Dict[int, int]] = None, transform_on_train: bool = False,
                 transform_on_eval: bool = True, transform_on_fantasize: bool = True, mc_samples: int = 128,
                 resample: bool = False, flip: bool = False, tau: float = 0.1, rng: str = "sobol",
                 philox_seed: int = 0, philox_offset: int = 0):
        super().__init__()
        if flip:
            raise NotImplementedError("flip=True only serves the arm/u2g estimators, dead code in the reference")
        self.transform_on_train = transform_on_train
        self.transform_on_eval = transform_on_eval
        self.transform_on_fantasize = transform_on_fantasize
        self._setup(integer_indices, integer_bounds, categorical_features, tau)
        self.mc_samples = mc_samples
        self.resample = resample
        self.flip = flip
        if rng not in ("sobol", "philox"):
            raise ValueError("rng must be 'sobol' or 'philox'")
        self.rng = rng
        self.philox_seed, self.philox_offset = int(philox_seed), int(philox_offset)

    # -- base samples --------------------------------------------------------------------------------------------------
    def _draw(self, n_cols: int, q: int, device, stream_col0: int) -> Tensor:
        if self.rng == "sobol":
            seed = torch.randint(0, 100000, (1,)).item()
            u = SobolEngine(q * n_cols, scramble=True, seed=seed).draw(self.mc_samples, dtype=torch.float64)
            return u.view(self.mc_samples, q, n_cols).to(device)
        if q != 1:
            raise NotImplementedError
        n_int = len(self._ints())
        n_cat = int(self.categorical_starts.shape[0])
        full = torch.empty(self.mc_samples, n_int + n_cat, dtype=torch.float64, device=device)
        with torch.cuda.device(device):
            capi.check(self.lib.prb_philox_uniform(self.philox_seed, self.philox_offset, self.mc_samples,
                                                   n_int + n_cat, full.data_ptr(), _stream(device)),
                       "prb_philox_uniform")
        if self.resample:
            self.philox_offset += 1
        return full[:, stream_col0:stream_col0 + n_cols].reshape(self.mc_samples, 1, n_cols).contiguous()

    def ensure_base_samples(self, q: int, device) -> None:
        n_int = len(self._ints())
        n_cat = int(self.categorical_starts.shape[0])
        if n_int > 0 and (not hasattr(self, "base_samples") or self.base_samples.shape[-2:] != (q, n_int)
                          or self.resample):
            self.register_buffer("base_samples", self._draw(n_int, q, device, 0))
        if n_cat > 0 and (not hasattr(self, "base_samples_categorical")
                          or self.base_samples_categorical.shape[-2] != q or self.resample):
            self.register_buffer("base_samples_categorical", self._draw(n_cat, q, device, n_int))

    def base_sample_ptrs(self):
        ui = self.base_samples if hasattr(self, "base_samples") else None
        uc = self.base_samples_categorical if hasattr(self, "base_samples_categorical") else None
        return ui, uc

    def transform(self, X: Tensor) -> Tensor:
        capi.require_cuda(X, "X")
        if X.ndim < 3 or X.shape[-3] != 1 or X.shape[-2] != 1:
            raise NotImplementedError("expected X of shape batch x 1 x q=1 x d")
        d = X.shape[-1]
        lead = X.shape[:-3]
        th = X.detach().to(torch.float64).reshape(-1, d).contiguous()
        self.ensure_base_samples(1, X.device)
        ui, uc = self.base_sample_ptrs()
        out = torch.empty(th.shape[0], self.mc_samples, d, dtype=torch.float64, device=X.device)
        sp = self._raw_space(d)
        with torch.cuda.device(X.device):
            capi.check(self.lib.prb_pr_sample(C.byref(sp), th.data_ptr(), th.shape[0], self.mc_samples,
                                              capi.ptr(ui), capi.ptr(uc), 0, 0, out.data_ptr(), None, None,
                                              _stream(X.device)), "prb_pr_sample")
        return out.reshape(*lead, self.mc_samples, 1, d)


class AnalyticProbabilisticReparameterizationInputTransform(_PRTransformBase):
    """``input.py:245-502``: enumeration of every discrete configuration (``all_discrete_options``), its probability
    under theta (``get_probs``) and the pasted inputs (``transform``)."""

    def __init__(self, dim: int, integer_indices: Optional[List[int]] = None, integer_bounds: Optional[Tensor] = None,
                 categorical_features: Optional[Dict[int, int]] = None, transform_on_train: bool = False,
                 transform_on_eval: bool = True, transform_on_fantasize: bool = True, tau: float = 0.1):
        super().__init__()
        self.transform_on_train = transform_on_train
        self.transform_on_eval = transform_on_eval
        self.transform_on_fantasize = transform_on_fantasize
        self._setup(integer_indices, integer_bounds, categorical_features, tau)
        self.dim = dim
        # cartesian product of the discrete options (:345-383); integers un-normalised, categoricals as indices
        opts = []
        if self.integer_indices is not None:
            ib = integer_bounds.detach().cpu()
            for i in range(ib.shape[-1]):
                opts.append(torch.arange(int(ib[0, i]), int(ib[1, i]) + 1, dtype=torch.long))
        cards = [int(e - s) for s, e in zip(self.categorical_starts.tolist(), self.categorical_ends.tolist())]
        for card in cards:
            opts.append(torch.arange(card, dtype=torch.long))
        codes = torch.cartesian_prod(*opts)
        if codes.ndim == 1:
            codes = codes.unsqueeze(-1)
        self.register_buffer("option_codes", codes.to(torch.int32).contiguous())
        n_disc = int(self.discrete_indices.shape[0])
        n_int = len(self._ints())
        table = torch.zeros(codes.shape[0], dim, dtype=torch.float64)
        for k, col in enumerate(self._ints()):
            table[:, col] = codes[:, k].to(torch.float64)
        for j, (s, card) in enumerate(zip(self.categorical_starts.tolist(), cards)):
            table[:, s:s + card] = torch.nn.functional.one_hot(codes[:, n_int + j], num_classes=card).to(table)
        self.register_buffer("all_discrete_options", table)
        assert n_disc == n_int + sum(cards)

    def get_probs(self, X: Tensor) -> Tensor:
        """UN-normalised ``X [b, 1, d]`` -> ``[b, n_discrete, 1]`` (``input.py:410-464``; values only -- gradients of the
        analytic objective come from ``prb_analytic_value_grad``)."""
        capi.require_cuda(X, "X")
        d = X.shape[-1]
        th = X.detach().to(torch.float64).reshape(-1, d).contiguous()
        n_opt = self.option_codes.shape[0]
        probs = torch.empty(th.shape[0], n_opt, dtype=torch.float64, device=X.device)
        sp = self._raw_space(d)
        codes = self.option_codes.to(X.device)
        with torch.cuda.device(X.device):
            capi.check(self.lib.prb_analytic_enumerate(C.byref(sp), th.data_ptr(), th.shape[0], codes.data_ptr(), n_opt,
                                                       None, probs.data_ptr(), _stream(X.device)),
                       "prb_analytic_enumerate")
        return probs.reshape(*X.shape[:-2], n_opt, X.shape[-2])

    def transform(self, X: Tensor) -> Tensor:
        """UN-normalised ``X [b, 1, 1, d]`` -> ``[b, n_discrete, 1, d]`` (``input.py:466-488``)."""
        capi.require_cuda(X, "X")
        d = X.shape[-1]
        lead = X.shape[:-3]
        th = X.detach().to(torch.float64).reshape(-1, d).contiguous()
        n_opt = self.option_codes.shape[0]
        out = torch.empty(th.shape[0], n_opt, d, dtype=torch.float64, device=X.device)
        sp = self._raw_space(d)
        codes = self.option_codes.to(X.device)
        with torch.cuda.device(X.device):
            capi.check(self.lib.prb_analytic_enumerate(C.byref(sp), th.data_ptr(), th.shape[0], codes.data_ptr(), n_opt,
                                                       out.data_ptr(), None, _stream(X.device)),
                       "prb_analytic_enumerate")
        return out.reshape(*lead, n_opt, 1, d)
