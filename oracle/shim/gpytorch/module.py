from torch.nn import Module as _Module


class Module(_Module):
    pass
