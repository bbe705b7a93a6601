"""Stand-in for botorch.models.model.Model (reference wrapper.py:14 imports the name only)."""
from torch.nn import Module


class Model(Module):
    pass
