"""Drop-in for the reference's env/hanoi_utils.py: hanoi_solver(state, goal_peg=2) -> int."""
from functools import lru_cache

import numpy as np
import torch

from .. import _lib
from ..engine import pack_state


@lru_cache(maxsize=None)
def hanoi_solver(state: tuple, goal_peg: int = 2) -> int:
    """Minimal number of moves from ``state`` to all disks on ``goal_peg`` (reference
    env/hanoi_utils.py:4-26), computed by the hmz_env_solver_distance kernel."""
    _lib.require_cuda()
    lib = _lib.load()
    n = len(state)
    if n > _lib.MAX_DISKS:
        raise ValueError(f"hanoi_solver: {n} disks exceed the packed-word limit of {_lib.MAX_DISKS}")
    w = torch.tensor([pack_state(state)], dtype=torch.int32, device="cuda")
    d = torch.empty(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.hmz_env_solver_distance(_lib.ptr(w), _lib.ptr(d), 1, n, goal_peg, _lib.current_stream()))
    return int(d.item())
