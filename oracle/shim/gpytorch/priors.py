class NormalPrior:
    def __init__(self, loc, scale, **kwargs):
        self.loc, self.scale = loc, scale
