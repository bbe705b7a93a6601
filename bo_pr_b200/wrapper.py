"""``discrete_mixed_bo/wrapper.py:19-32``: acquisition-function wrapper base (holds ``acq_function``, proxies
``X_baseline`` and ``model``)."""
from __future__ import annotations

from typing import Optional

from torch import Tensor
from torch.nn import Module

from .acquisition import AcquisitionFunction


class AcquisitionFunctionWrapper(AcquisitionFunction):
    r"""Abstract acquisition wrapper."""

    def __init__(self, acq_function: AcquisitionFunction) -> None:
        Module.__init__(self)
        self.acq_function = acq_function
        self.X_pending = None

    @property
    def X_baseline(self) -> Optional[Tensor]:
        return getattr(self.acq_function, "X_baseline", None)

    @property
    def model(self):
        return self.acq_function.model
