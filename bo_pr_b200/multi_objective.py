"""Box decomposition of the non-dominated region for analytic EHVI.

The reference builds ``FastNondominatedPartitioning(ref_point=ref_point, Y=train_Y)`` (``experiment_utils.py:395-402``,
[3P] ``botorch.utils.multi_objective.box_decompositions``) and hands it to ``ExpectedHypervolumeImprovement``, which reads
``get_hypercell_bounds()``.  EHVI is an integral over the non-dominated region, so its value does not depend on WHICH
disjoint boxes tile that region; this module produces such a tiling with a dimension sweep (host side, once per BO
iteration, tens of Pareto points):

    slab j of the last objective, y_m in [v_j, v_{j+1}) between consecutive Pareto values (v_0 = ref_m, last slab up to
    +inf): a point of the slab is dominated exactly by the Pareto points with p_m >= v_{j+1}, so the slab's non-dominated
    part is (non-dominated region of those points' first m-1 coordinates) x [v_j, v_{j+1}); recurse down to one objective,
    where the region is [max p, +inf).

All objectives are maximised.  The arithmetic of EHVI itself runs in ``multi_ehvi_epilogue_kernel``.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch
from torch import Tensor

Box = Tuple[Tuple[float, ...], Tuple[float, ...]]


def is_non_dominated(Y: Tensor) -> Tensor:
    """Mask of the rows of ``Y [n, m]`` that no other row dominates (>= everywhere, > somewhere); of exact duplicates the
    first is kept."""
    n = Y.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.bool, device=Y.device)
    ge = (Y.unsqueeze(0) >= Y.unsqueeze(1)).all(-1)  # ge[i, j]: row j >= row i everywhere
    gt = (Y.unsqueeze(0) > Y.unsqueeze(1)).any(-1)
    dominated = (ge & gt).any(-1)
    eq = (Y.unsqueeze(0) == Y.unsqueeze(1)).all(-1)
    earlier_dup = torch.tril(eq, diagonal=-1).any(-1)
    return ~dominated & ~earlier_dup


def _pareto(points: List[Tuple[float, ...]]) -> List[Tuple[float, ...]]:
    out = []
    for i, p in enumerate(points):
        dom = False
        for j, q in enumerate(points):
            if j != i and all(a >= b for a, b in zip(q, p)) and (any(a > b for a, b in zip(q, p)) or j < i):
                dom = True
                break
        if not dom:
            out.append(p)
    return out


def _cells(points: List[Tuple[float, ...]], ref: Tuple[float, ...]) -> List[Box]:
    m = len(ref)
    if m == 1:
        top = max([p[0] for p in points] + [ref[0]])
        return [((top,), (math.inf,))]
    cells: List[Box] = []
    z = ref[-1]
    for v in sorted({p[-1] for p in points}) + [math.inf]:
        sub = _pareto([p[:-1] for p in points if p[-1] >= v]) if v != math.inf else []
        for lo, hi in _cells(sub, ref[:-1]):
            cells.append((lo + (z,), hi + (v,)))
        z = v
    return cells


def _hypervolume(points: List[Tuple[float, ...]], ref: Tuple[float, ...]) -> float:
    m = len(ref)
    if not points:
        return 0.0
    if m == 1:
        return max(p[0] for p in points) - ref[0]
    hv, z = 0.0, ref[-1]
    for v in sorted({p[-1] for p in points}):
        hv += (v - z) * _hypervolume(_pareto([p[:-1] for p in points if p[-1] >= v]), ref[:-1])
        z = v
    return hv


class FastNondominatedPartitioning:
    """Surface of ``botorch.utils.multi_objective.box_decompositions.non_dominated.FastNondominatedPartitioning`` that the
    reference touches: constructor ``(ref_point, Y)``, ``pareto_Y``, ``get_hypercell_bounds()`` (``[2, n_cells, m]``, lower
    corners then upper corners, ``+inf`` where a cell is unbounded), ``compute_hypervolume()``, ``update(Y)``."""

    def __init__(self, ref_point: Tensor, Y: Tensor = None) -> None:
        self.ref_point = torch.as_tensor(ref_point).detach().reshape(-1)
        self.num_outcomes = int(self.ref_point.numel())
        self._Y = None
        self.update(Y if Y is not None else self.ref_point.new_zeros(0, self.num_outcomes), reset=True)

    def update(self, Y: Tensor, reset: bool = False) -> None:
        Y = torch.as_tensor(Y).detach().reshape(-1, self.num_outcomes).to(self.ref_point)
        if not reset and self._Y is not None:
            Y = torch.cat([self._Y, Y], dim=0)
        self._Y = Y
        better = Y[(Y > self.ref_point).all(-1)]  # only points strictly better than the reference point can dominate
        self.pareto_Y = better[is_non_dominated(better)]
        ref = tuple(float(v) for v in self.ref_point.cpu().tolist())
        pts = [tuple(float(v) for v in row) for row in self.pareto_Y.cpu().tolist()]
        boxes = _cells(pts, ref)
        lo = torch.tensor([b[0] for b in boxes], dtype=self.ref_point.dtype)
        hi = torch.tensor([b[1] for b in boxes], dtype=self.ref_point.dtype)
        self._bounds = torch.stack([lo, hi]).to(self.ref_point.device)
        self._points, self._ref = pts, ref

    def get_hypercell_bounds(self) -> Tensor:
        return self._bounds

    def compute_hypervolume(self) -> Tensor:
        return torch.tensor(_hypervolume(self._points, self._ref), dtype=self.ref_point.dtype,
                            device=self.ref_point.device)


NondominatedPartitioning = FastNondominatedPartitioning
