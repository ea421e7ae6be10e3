"""GPU: batched acting evaluation (acting.get_results, the lesion harness of
acting_experiments/acting_ablations.py:72-128) against the oracle port playing the same episodes one at a
time with the same start states and the same sampling uniforms."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu


def _oracle_episode(n, max_steps, sd, start_idx, n_sims, temperature, uniforms):
    env = port.PortHanoi(n, max_steps, init_state_idx=start_idx)
    obs = env.reset()
    net, search = port.PortNet(sd), port.PortSearch(0.8, n_sims)  # fresh MinMaxStats per episode
    start = tuple(env.current_state())
    steps, illegal, done = 0, 0, False
    while not done:
        a, _, _, _, _ = port.run_mcts_port(obs, net, search, temperature, False, alpha=0.0, u=float(uniforms[steps]))
        obs, _, done, ill = env.step(a)
        illegal += int(ill)
        steps += 1
    return steps - port.hanoi_solver(start), steps, illegal


@pytest.mark.parametrize("heads", [(), ("policy_net", "value_net")])
def test_get_results_matches_oracle_episodes(heads):
    from muzero_hanoi_b200 import _lib, acting
    from muzero_hanoi_b200.engine import PackedWeights

    n, max_steps, B, n_sims, T = 3, 30, 24, 12, 0.0
    sd = port.make_weights(n, 11)
    if heads:
        sd = port.lesion_weights(sd, heads, seed=5)
    rng = np.random.default_rng(3)
    starts = rng.integers(0, 26, B)  # non-goal states (index 26 = goal (2,2,2))
    uniforms = rng.random((max_steps, B))
    data, det = acting.get_results(PackedWeights(sd, n, _lib.MODE_FP32), n, max_steps, B, [n_sims], T, start_indices=starts,
                                   uniforms=uniforms, return_details=True)
    torch.cuda.synchronize()
    ref = [_oracle_episode(n, max_steps, sd, int(starts[g]), n_sims, T, uniforms[:, g]) for g in range(B)]
    err = np.array([r[0] for r in ref])
    same = (det[0]["errors"] == err) & (det[0]["steps"] == np.array([r[1] for r in ref]))
    # the float32 network kernels agree with torch to ~1e-6, so a near-tie in one search can move one episode
    assert same.mean() >= 0.9, (det[0]["errors"], err)
    if same.all():
        assert data == [[n_sims, float(err.sum()) / B]]
        assert np.array_equal(det[0]["illegal_moves"], np.array([r[2] for r in ref]))
    assert (det[0]["steps"] >= 1).all() and (det[0]["steps"] <= max_steps).all()
    assert np.array_equal(det[0]["min_moves"], [port.hanoi_solver(port.index_to_state(int(i), n)) for i in starts])


def test_get_results_shared_minmax_is_sequentially_consistent_with_the_reference():
    """shared_minmax=True: ONE MinMaxStats threads through every episode of the run (MCTS/mcts.py:23,
    acting_ablations.py:343), so an episode's searches are normalised with the extrema of all earlier episodes.  The
    oracle plays the same episodes in the same order on one PortSearch object; also checks that the option matters
    (the per-episode-stats default gives different bounds)."""
    from muzero_hanoi_b200 import _lib, acting
    from muzero_hanoi_b200.engine import PackedWeights

    n, max_steps, B, n_sims, T = 3, 30, 10, 12, 0.0
    sd = port.make_weights(n, 11)
    rng = np.random.default_rng(8)
    starts = rng.integers(0, 26, B)
    uniforms = rng.random((max_steps, B))
    _, det = acting.get_results(PackedWeights(sd, n, _lib.MODE_FP32), n, max_steps, B, [n_sims], T, start_indices=starts,
                                uniforms=uniforms, return_details=True, shared_minmax=True)
    net, search = port.PortNet(sd), port.PortSearch(0.8, n_sims)  # ONE MinMaxStats for all episodes
    ref_steps = []
    for g in range(B):
        env = port.PortHanoi(n, max_steps, init_state_idx=int(starts[g]))
        obs, done, steps = env.reset(), False, 0
        while not done:
            a, _, _, _, _ = port.run_mcts_port(obs, net, search, T, False, alpha=0.0, u=float(uniforms[steps, g]))
            obs, _, done, _ = env.step(a)
            steps += 1
        ref_steps.append(steps)
    same = det[0]["steps"] == np.array(ref_steps)
    assert same.mean() >= 0.9, (det[0]["steps"], ref_steps)
    assert np.array_equal(det[0]["errors"], det[0]["steps"] - det[0]["min_moves"])


def test_compute_n_step_returns_dropin_matches_oracle():
    from muzero_hanoi_b200.utils import compute_n_step_returns

    rw = [0, -100 / 1000, 0, 0, -100 / 1000, 100]
    q = [1.25, -3.5, 0.75, 10.0, 44.0, 80.5]
    assert compute_n_step_returns(rw, q, 3, 0.8) == port.n_step_returns(rw, q, 3, 0.8)
    with pytest.raises(ValueError):
        compute_n_step_returns([1.0], [0.0], 3, 0.8)
