"""Stand-in for botorch.acquisition.acquisition (surface used by reference wrapper.py:13, probabilistic_reparameterization.py:17)."""
from torch.nn import Module


class AcquisitionFunction(Module):
    def __init__(self, model=None):
        super().__init__()
        if model is not None:
            self.model = model

    def set_X_pending(self, X_pending=None):
        self.X_pending = X_pending
