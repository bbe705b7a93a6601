class Interval:
    def __init__(self, lower_bound, upper_bound, **kwargs):
        self.lower_bound, self.upper_bound = lower_bound, upper_bound


class GreaterThan(Interval):
    def __init__(self, lower_bound, **kwargs):
        super().__init__(lower_bound, float("inf"))
