"""Batched acting evaluation: ``acting_ablations.get_results`` / ``ablate_networks`` of the reference
(acting_experiments/acting_ablations.py:29-45, 72-128) with every episode of a setting run as one game
of a lock-step batch on the GPU.

Differences from the sequential reference, both forced by running the episodes in parallel: by default every
episode owns its MinMaxStats (the reference threads ONE stats object through all episodes of a run,
MCTS/mcts.py:23 and acting_ablations.py:343, so later episodes see the extrema of earlier ones), and the random
draws are inputs (``uniforms`` per move) or on-device Philox streams instead of the process-global NumPy state.
``shared_minmax=True`` reproduces the reference's experiment exactly in that respect: the episodes of a setting then
run ONE AFTER ANOTHER through a single search slot whose (min, max) persists from episode to episode — sequentially
consistent with the reference, and as serial as the reference is.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, current_stream, ptr
from .engine import BatchedMCTS, PackedWeights, VecHanoi


def ablate_networks(reset_latent_policy, reset_latent_values, reset_latent_rwds, networks):
    """acting_ablations.py:29-45: re-initialise the chosen heads with ``networks.reset_param``."""
    if reset_latent_policy:
        networks.policy_net.apply(networks.reset_param)
    if reset_latent_values:
        networks.value_net.apply(networks.reset_param)
    if reset_latent_rwds:
        networks.rwd_net.apply(networks.reset_param)
    return networks


def get_results(weights: PackedWeights, N, max_steps, episode, n_mcts_simulations_range, temperature, *, start_idx=None,
                discount=0.8, root_dirichlet_alpha=0.0, seed=0, uniforms=None, start_indices=None, device="cuda",
                latent_dtype=None, return_details=False, shared_minmax=False):
    """``[[n_simulations, mean(steps - hanoi_solver(start))], ...]`` (acting_ablations.py:72-128) over
    ``episode`` parallel episodes per simulation budget.

    start_idx      index into env.states to reset to (``env.init_state_idx``); None = ``random_reset``
    start_indices  explicit start state index per episode (parity tests); overrides start_idx
    uniforms       float64 [moves, episode]: the ``np.random.choice`` draw of each move (MCTS/mcts.py:120);
                   drawn on device from Philox streams keyed by ``seed`` when omitted
    """
    _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device(device)
    B = int(episode)
    ldt = (_lib.LATENT_BF16 if weights.mode == _lib.MODE_BF16 else _lib.LATENT_F32) if latent_dtype is None else latent_dtype
    data, details = [], []
    if shared_minmax:
        return _get_results_sequential(weights, N, max_steps, B, n_mcts_simulations_range, temperature, start_idx, discount,
                                       root_dirichlet_alpha, seed, uniforms, start_indices, dev, ldt, return_details)
    for n_sims in n_mcts_simulations_range:
        env = VecHanoi(N, max_steps, B, dev, init_state_idx=start_idx or 0)
        if start_indices is not None:
            env.set_state_indices(np.asarray(start_indices, dtype=np.int32))
        elif start_idx is not None:
            env.reset()
        else:
            env.random_reset(seed=seed, counter=int(n_sims))
        min_moves = env.solver_distance()  # hanoi_solver(tuple(env.current_state())), :96-98
        mcts = BatchedMCTS(discount, root_dirichlet_alpha, int(n_sims), B, dev, latent_dtype=ldt)
        steps = torch.zeros(B, dtype=torch.int32, device=dev)
        illegal = torch.zeros(B, dtype=torch.int32, device=dev)
        errors = torch.empty(B, dtype=torch.int32, device=dev)
        u_dev = torch.empty(B, dtype=torch.float64, device=dev)
        noise = torch.empty(B, 6, dtype=torch.float64, device=dev) if root_dirichlet_alpha > 0 else None
        for move in range(max_steps):  # every episode ends by max_steps (env/hanoi.py:77-80)
            if uniforms is not None:
                u_dev.copy_(torch.as_tensor(uniforms[move], dtype=torch.float64))
            else:
                check(lib.hmz_rng_uniform(ptr(u_dev), B, seed, (int(n_sims) << 20) | move, 0, current_stream()))
            if noise is not None:
                check(lib.hmz_rng_dirichlet(ptr(noise), B, float(root_dirichlet_alpha), seed, (int(n_sims) << 20) | move, 0,
                                            current_stream()))
            action, _, _, _ = mcts.run_mcts(weights, words=env.words, temperature=temperature, deterministic=False,
                                            noise=noise, uniforms=u_dev)
            _, _, flags = env.step(action, want_obs=False)
            check(lib.hmz_eval_track(ptr(flags), move, B, ptr(steps), ptr(illegal), current_stream()))
            if move % 8 == 7 and bool((steps > 0).all()):  # all first episodes over: stop early
                break
        check(lib.hmz_eval_errors(ptr(steps), ptr(min_moves), B, ptr(errors), current_stream()))
        err = errors.cpu().numpy()
        data.append([n_sims, float(err.sum()) / len(err)])  # sum(errors) / len(errors), :123
        details.append(dict(n_simulations=n_sims, steps=steps.cpu().numpy(), errors=err, illegal_moves=illegal.cpu().numpy(),
                            min_moves=min_moves.cpu().numpy()))
    return (data, details) if return_details else data


def _get_results_sequential(weights, N, max_steps, B, n_range, temperature, start_idx, discount, alpha, seed, uniforms,
                            start_indices, dev, ldt, return_details):
    """``shared_minmax=True``: acting_ablations.get_results exactly as the reference sequences it — one MCTS object per
    simulation budget (acting_ablations.py:76-79 mutates ``n_simulations`` on the SAME object, so the MinMaxStats even
    carries over from one budget to the next), episodes played one after another (:84-123)."""
    lib = _lib.load()
    data, details = [], []
    cap = max(int(n) for n in n_range)
    mcts = BatchedMCTS(discount, alpha, cap, 1, dev, latent_dtype=ldt)  # the one MinMaxStats of the run (MCTS/mcts.py:23)
    u_dev = torch.empty(1, dtype=torch.float64, device=dev)
    noise = torch.empty(1, 6, dtype=torch.float64, device=dev) if alpha > 0 else None
    for n_sims in n_range:
        mcts.n_simulations = int(n_sims)
        steps_all, err_all, ill_all, min_all = [], [], [], []
        for ep in range(B):
            env = VecHanoi(N, max_steps, 1, dev, init_state_idx=start_idx or 0, auto_reset=False)
            if start_indices is not None:
                env.set_state_indices(np.asarray([start_indices[ep]], dtype=np.int32))
            elif start_idx is not None:
                env.reset()
            else:
                env.random_reset(seed=seed, counter=(int(n_sims) << 20) | ep)  # (its own stream: the start states differ from the batched form's)
            min_moves = int(env.solver_distance().item())
            steps = illegal = 0
            done = False
            while not done:
                ctr = (int(n_sims) << 20) | steps  # item = the episode: the draws of the batched form's game `ep`
                if uniforms is not None:
                    u_dev.fill_(float(uniforms[steps][ep]))
                else:
                    check(lib.hmz_rng_uniform(ptr(u_dev), 1, seed, ctr, ep, current_stream()))
                if noise is not None:
                    check(lib.hmz_rng_dirichlet(ptr(noise), 1, float(alpha), seed, ctr, ep, current_stream()))
                action, _, _, _ = mcts.run_mcts(weights, words=env.words, temperature=temperature, deterministic=False, noise=noise,
                                                uniforms=u_dev)
                _, _, flags = env.step(action, want_obs=False)
                f = int(flags.item())
                illegal += int(bool(f & _lib.FLAG_ILLEGAL))
                steps += 1
                done = bool(f & _lib.FLAG_DONE)
            steps_all.append(steps), err_all.append(steps - min_moves), ill_all.append(illegal), min_all.append(min_moves)
        data.append([n_sims, float(sum(err_all)) / len(err_all)])
        details.append(dict(n_simulations=n_sims, steps=np.array(steps_all, np.int32), errors=np.array(err_all, np.int32),
                            illegal_moves=np.array(ill_all, np.int32), min_moves=np.array(min_all, np.int32)))
    return (data, details) if return_details else data
