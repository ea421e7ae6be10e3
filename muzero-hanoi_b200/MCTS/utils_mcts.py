"""Drop-in for the reference's MCTS/utils_mcts.py."""


class MinMaxStats(object):
    """Running (min, max) of the tree values (reference MCTS/utils_mcts.py:1-16).  On an MCTS
    object this is the host mirror of the device-resident pair the kernels update; the crossed
    constructor arguments of the reference are kept as they are."""

    def __init__(self, min_value_bound=None, max_value_bound=None):
        self.maximum = min_value_bound if min_value_bound else -float("inf")
        self.minimum = max_value_bound if max_value_bound else float("inf")

    def update(self, value):
        self.maximum = max(self.maximum, value)
        self.minimum = min(self.minimum, value)

    def normalize(self, value):
        if self.maximum > self.minimum:
            return (value - self.minimum) / (self.maximum - self.minimum)
        return value
