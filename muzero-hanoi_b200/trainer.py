"""Batched counterpart of the reference's ``Muzero`` class (Muzero.py): self-play, episode post-processing,
replay and the learner step, all on the device, tied together in the shape of ``Muzero.training_loop``
(Muzero.py:80-150).

Differences from the sequential reference, all forced by playing B games at once:
  * one "loop" plays ``moves_per_loop`` moves of all B games instead of ``n_ep_x_loop`` whole episodes; the episodes
    that finish during a loop are post-processed and (if solved, Muzero.py:98) stored;
  * the temperature schedule (utils.adjust_temperature) is driven by the number of finished episodes per game slot;
  * acting uses the weights of the last completed update (repacked for the acting kernels once per loop).
Same hyper-parameters and defaults as training_main.TrainingConfig (training_main.py:17-36).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import PackedWeights, SelfPlay
from .learner import Learner
from .replay import ReplayRing
from .utils import adjust_temperature


class BatchedMuzero:
    def __init__(self, state_dict, N, max_steps, n_games, discount=0.8, dirichlet_alpha=0.25, n_mcts_simulations=25,
                 unroll_n_steps=5, batch_s=256, n_TD_step=10, lr=0.002, buffer_size=50000, priority_replay=True,
                 n_update_x_loop=1, moves_per_loop=8, acting_mode=_lib.MODE_FP32, seed=0, device="cuda"):
        self.N, self.max_steps, self.B = int(N), int(max_steps), int(n_games)
        self.discount, self.alpha, self.S = discount, dirichlet_alpha, int(n_mcts_simulations)
        self.unroll, self.batch_s, self.n_step = int(unroll_n_steps), int(batch_s), int(n_TD_step)
        self.priority_replay, self.n_update_x_loop, self.moves_per_loop = priority_replay, int(n_update_x_loop), int(moves_per_loop)
        self.acting_mode, self.seed, self.device = acting_mode, int(seed), device
        self.learner = Learner(state_dict, N, unroll_n_steps, lr=lr, device=device)
        self.buffer = ReplayRing(buffer_size, unroll_n_steps, 3 * N, 6, device)
        self.episodes_done = 0
        self._selfplay = None
        self._temperature = None
        self._weights_dirty = False

    def _latent_dtype(self):
        return _lib.LATENT_BF16 if self.acting_mode == _lib.MODE_BF16 else _lib.LATENT_F32

    def _refresh_actor(self, temperature):
        """New acting weights (and a new temperature) take effect; games in flight and their episode store carry on.
        The weights are repacked only after an update has changed them."""
        if self._selfplay is None:
            w = PackedWeights({k: v for k, v in self.learner.state_dict().items()}, self.N, self.acting_mode, self.device)
            self._selfplay = SelfPlay(self.N, self.max_steps, self.B, self.S, w, self.discount, self.alpha, temperature=temperature,
                                      seed=self.seed, device=self.device, latent_dtype=self._latent_dtype(), episodes=True)
        else:
            if self._weights_dirty:
                self._selfplay.weights = PackedWeights({k: v for k, v in self.learner.state_dict().items()}, self.N, self.acting_mode,
                                                       self.device)
            self._selfplay.temperature = float(temperature)
        self._weights_dirty = False
        self._temperature = temperature

    def play(self, n_moves):
        """n_moves moves of every game; finished episodes are post-processed and, if solved, stored.
        Returns (episodes finished, mean length of those episodes, transitions stored)."""
        sp, st = self._selfplay, self._selfplay.episodes
        stored = 0
        counts = torch.zeros(2, dtype=torch.int64, device=st.ep_len.device)  # episodes finished, sum of their lengths
        for _ in range(n_moves):
            sp.move()
            st.post_process(self.n_step, self.discount)
            stored += self.buffer.add_episodes(st, temperature=self._temperature, only_solved=True)  # the one sync per move
            lens = st.ep_len
            counts[0] += (lens > 0).sum()
            counts[1] += lens.sum()
        finished, length_sum = (int(x) for x in counts.tolist())  # read back once per call
        self.episodes_done += finished
        return finished, (length_sum / finished if finished else float("nan")), stored

    def update(self):
        """One Muzero._update on a batch drawn from the replay ring (Muzero.py:104-127)."""
        if self.priority_replay:
            states, rwds, actions, pi_probs, returns, indx, w = self.buffer.priority_sample(self.batch_s)
        else:
            states, rwds, actions, pi_probs, returns = self.buffer.uniform_sample(self.batch_s)
            indx, w = None, None
        new_p, v_loss, r_loss, p_loss = self.learner.update(states, rwds, actions, pi_probs, returns, w)
        self._weights_dirty = True
        self.buffer.update_priorities(indx, new_p)
        return v_loss, r_loss, p_loss

    def training_loop(self, n_loops, min_replay_size, print_acc=None, log=print):
        """-> list of (loop, mean episode length of the loop, value loss, reward loss, policy loss)."""
        history = []
        for n in range(1, n_loops):
            self._refresh_actor(adjust_temperature(self.episodes_done // max(1, self.B)))
            finished, mean_len, _ = self.play(self.moves_per_loop)
            losses = (float("nan"),) * 3
            if len(self.buffer) > min_replay_size:
                for _ in range(self.n_update_x_loop):
                    losses = self.update()
            history.append((n, mean_len, *losses))
            if print_acc and n % print_acc == 0:
                log("Loop %d | episodes %d | steps %.3f | V %.3f | rwd %.3f | Pi %.3f" % (n, self.episodes_done, mean_len, *losses))
        return history
