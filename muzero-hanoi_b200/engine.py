"""Batched engines over libhmz.so: VecHanoi (env), PackedWeights, BatchedMCTS, SelfPlay.

These are the batched additions SURVEY.md §8b names; the reference-shaped classes in
``env/``, ``MCTS/`` and ``networks.py`` are thin B=1 views over them.  torch is used for device
memory and streams only; every computation is a libhmz kernel.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import check, current_stream, ptr

MOVES = ((0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1))  # env/hanoi.py:39-41


# ----------------------------------------------------------------------- packing helpers
def pack_state(state) -> int:
    """tuple(peg of disk d) -> env word state bits (2 bits per disk, disk 0 lowest)."""
    w = 0
    for d, p in enumerate(state):
        w |= (int(p) & 3) << (2 * d)
    return w


def unpack_state(word: int, n_disks: int) -> tuple:
    return tuple((int(word) >> (2 * d)) & 3 for d in range(n_disks))


def state_index(state) -> int:
    """Index into itertools.product(range(3), repeat=N) (disk 0 most significant)."""
    idx = 0
    for p in state:
        idx = idx * 3 + int(p)
    return idx


def index_state(idx: int, n_disks: int) -> tuple:
    out = [0] * n_disks
    for d in range(n_disks - 1, -1, -1):
        out[d] = idx % 3
        idx //= 3
    return tuple(out)


def check_env_shape(n_disks: int, max_steps: int):
    if not 1 <= n_disks <= _lib.MAX_DISKS:
        raise ValueError(f"N={n_disks} unsupported: the packed env word holds 1..{_lib.MAX_DISKS} disks")
    if not 1 <= max_steps < (1 << (32 - 2 * n_disks)):
        raise ValueError(f"max_steps={max_steps} does not fit the {32 - 2 * n_disks} counter bits of the env word")


class VecHanoi:
    """B Tower-of-Hanoi environments stepped in lock-step on the GPU.

    State lives in one int32 tensor ``words`` (bit layout in include/hmz.h).  Semantics are
    those of TowersOfHanoi.step (reference env/hanoi.py:47-84) per env, bit-exact."""

    def __init__(self, N, max_steps, B, device="cuda", init_state_idx=0, goal_peg=2, auto_reset=True):
        _lib.require_cuda()
        check_env_shape(N, max_steps)
        self.lib = _lib.load()
        self.discs, self.max_steps, self.B, self.goal_peg = N, max_steps, int(B), goal_peg
        self.device = torch.device(device)
        self.init_state_idx = init_state_idx
        self.auto_reset = bool(auto_reset)
        self.words = torch.zeros(self.B, dtype=torch.int32, device=self.device)
        self.rewards = torch.empty(self.B, dtype=torch.float32, device=self.device)
        self.flags = torch.empty(self.B, dtype=torch.uint8, device=self.device)
        self.obs_words = torch.empty(self.B, dtype=torch.int32, device=self.device)
        self.actions = torch.empty(self.B, dtype=torch.uint8, device=self.device)
        self._counters = torch.zeros(4, dtype=torch.int64, device=self.device)

    @property
    def reset_word(self) -> int:
        return pack_state(index_state(self.init_state_idx, self.discs))

    @property
    def goal_word(self) -> int:
        return pack_state((self.goal_peg,) * self.discs)

    def reset(self):
        check(self.lib.hmz_env_reset(ptr(self.words), self.B, self.reset_word, current_stream()))
        return self.words

    def set_state_indices(self, indices):
        idx = torch.as_tensor(indices, dtype=torch.int32, device=self.device).contiguous()
        assert idx.numel() == self.B
        check(self.lib.hmz_env_from_index(ptr(idx), ptr(self.words), self.B, self.discs, current_stream()))
        return self.words

    def state_indices(self):
        out = torch.empty(self.B, dtype=torch.int32, device=self.device)
        check(self.lib.hmz_env_to_index(ptr(self.words), ptr(out), self.B, self.discs, current_stream()))
        return out

    def random_reset(self, seed=0, counter=0):
        check(self.lib.hmz_env_random_reset(ptr(self.words), self.B, self.discs, self.goal_peg, seed, counter,
                                            current_stream()))
        return self.words

    def step(self, actions, want_obs=True):
        """actions: uint8 tensor [B] on the device.  Returns (obs_words|None, rewards, flags)."""
        a = actions if actions.dtype == torch.uint8 else actions.to(torch.uint8)
        check(self.lib.hmz_env_step(ptr(self.words), ptr(a), ptr(self.rewards), ptr(self.flags),
                                    ptr(self.obs_words) if want_obs else None, self.B, self.discs, self.max_steps,
                                    self.goal_peg, int(self.auto_reset), self.reset_word, current_stream()))
        return (self.obs_words if want_obs else None), self.rewards, self.flags

    def step_random(self, seed=0, step_index=0):
        check(self.lib.hmz_env_step_random(ptr(self.words), ptr(self.actions), ptr(self.rewards), ptr(self.flags),
                                           self.B, self.discs, self.max_steps, self.goal_peg, self.reset_word, seed,
                                           step_index, current_stream()))
        return self.actions, self.rewards, self.flags

    def rollout_random(self, n_steps, seed=0, step_index=0):
        """n_steps fused random-legal-move steps; returns the int64[4] device counters
        (steps, goals, truncations, xor checksum), accumulated since construction."""
        check(self.lib.hmz_env_rollout_random(ptr(self.words), self.B, self.discs, self.max_steps, self.goal_peg,
                                              self.reset_word, n_steps, seed, step_index, ptr(self._counters),
                                              current_stream()))
        return self._counters

    def legal_mask(self):
        out = torch.empty(self.B, dtype=torch.uint8, device=self.device)
        check(self.lib.hmz_env_legal_mask(ptr(self.words), ptr(out), self.B, self.discs, current_stream()))
        return out

    def onehot(self, words=None):
        w = self.words if words is None else words
        out = torch.empty(w.numel(), 3 * self.discs, dtype=torch.float32, device=self.device)
        check(self.lib.hmz_env_onehot(ptr(w), ptr(out), w.numel(), self.discs, current_stream()))
        return out

    def solver_distance(self, words=None):
        w = self.words if words is None else words
        out = torch.empty(w.numel(), dtype=torch.int32, device=self.device)
        check(self.lib.hmz_env_solver_distance(ptr(w), ptr(out), w.numel(), self.discs, self.goal_peg,
                                               current_stream()))
        return out

    def states(self):
        """Host view: (state bits [B] uint32, step counters [B])."""
        w = self.words.cpu().numpy().view(np.uint32)
        shift = 2 * self.discs
        return w & ((1 << shift) - 1), w >> shift
