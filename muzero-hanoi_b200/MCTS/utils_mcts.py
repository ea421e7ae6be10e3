"""Host mirror of the search's running value bounds (the reference's MCTS/utils_mcts.py surface).

On an ``MCTS`` object this pair shadows the device-resident ``minmax[search][2]`` array that the tree
kernels update (``hmz_search_minmax_reset`` / ``search_backup_select``): ``MCTS.run_mcts`` uploads it before
a search and reads it back afterwards, so callers that inspect or replace ``mcts.min_max_stats`` see what
they would see with the reference.  Semantics restated from MCTS/utils_mcts.py:1-16:

* start at (+inf, -inf); the two optional constructor bounds are accepted in the reference's order, where
  the first one seeds ``maximum`` and the second one ``minimum`` (no caller passes them);
* ``update(x)`` widens the interval; ``normalize(x)`` maps onto [0, 1] only once the interval is non-empty.
"""
import math


class MinMaxStats:
    __slots__ = ("maximum", "minimum")

    def __init__(self, min_value_bound=None, max_value_bound=None):
        lo_seed, hi_seed = max_value_bound, min_value_bound  # the reference's (crossed) assignment
        self.minimum = lo_seed if lo_seed else math.inf
        self.maximum = hi_seed if hi_seed else -math.inf

    def update(self, value):
        if value > self.maximum:
            self.maximum = value
        if value < self.minimum:
            self.minimum = value

    def normalize(self, value):
        span = self.maximum - self.minimum
        if not self.maximum > self.minimum:
            return value
        return (value - self.minimum) / span

    def as_pair(self):
        """(minimum, maximum) in the order of the device array."""
        return self.minimum, self.maximum
