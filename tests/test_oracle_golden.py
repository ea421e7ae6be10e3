"""CPU: the oracle port (oracle/port.py) against the golden fixtures generated from the
unmodified reference (tests/golden/, made by oracle/gen_golden.py)."""
import numpy as np
import pytest

from oracle import port

REWARDS = {0: 0, 1: 100, 2: -100 / 1000}
SEARCH_FIXTURES = ["n3_s50_noise_t1", "n3_s25_nonoise_t0", "n3_s25_det", "n4_s200_lesion_t0", "n5_s100_noise_t05"]


@pytest.mark.parametrize("n", [3, 4, 5, 7])
def test_env_transitions_exhaustive(golden, n):
    g = golden("env_tables.npz")
    max_steps = int(g["max_steps"])
    goal = (2,) * n
    for vi, c0 in enumerate(g["counter_before"]):
        for si in range(3 ** n):
            st = port.index_to_state(si, n)
            for a in range(6):
                moved, stored, ctr, r, d, ill, rc = port.step_state(st, int(c0), a, max_steps, goal)
                assert port.state_to_index(moved) == g[f"n{n}_obs_idx"][vi, si, a]
                assert port.state_to_index(stored) == g[f"n{n}_stored_idx"][vi, si, a]
                assert r == REWARDS[int(g[f"n{n}_reward_code"][vi, si, a])]
                assert (d, ill, ctr, rc) == (bool(g[f"n{n}_done"][vi, si, a]), bool(g[f"n{n}_illegal"][vi, si, a]),
                                             int(g[f"n{n}_counter_after"][vi, si, a]),
                                             bool(g[f"n{n}_reset_check"][vi, si, a]))


def test_env_transitions_n10_sampled(golden):
    g = golden("env_tables.npz")
    rng = np.random.default_rng(0)
    for si in rng.integers(0, 3 ** 10, 3000):
        st = port.index_to_state(int(si), 10)
        assert port.hanoi_solver(st) == g["n10_solver"][si] and port.legal_mask(st) == g["n10_legal"][si]
        for a in range(6):
            moved, stored, ctr, r, d, ill, _ = port.step_state(st, 17, a, 200, (2,) * 10)
            assert port.state_to_index(moved) == g["n10_obs_idx"][0, si, a]
            assert r == REWARDS[int(g["n10_reward_code"][0, si, a])] and ill == bool(g["n10_illegal"][0, si, a])


@pytest.mark.parametrize("n", [3, 4, 5, 7])
def test_solver_and_legal_masks(golden, n):
    g = golden("env_tables.npz")
    for si in range(3 ** n):
        st = port.index_to_state(si, n)
        assert port.hanoi_solver(st) == g[f"n{n}_solver"][si]
        assert port.legal_mask(st) == g[f"n{n}_legal"][si]
        assert port.packed_to_state(port.state_to_packed(st), n) == st


def test_known_answers():
    # acting_ablations.py:53-60 / generate_all_figures.py:78 (OPTIMAL_MOVES) of the reference
    assert [port.hanoi_solver(s) for s in ((2, 2, 0), (0, 0, 2), (1, 2, 2), (0, 0, 0))] == [7, 3, 1, 7]
    assert port.MOVES == ((0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1))
    # sample rows quoted in SURVEY.md §8c
    assert port.step_state((0, 2, 2), 0, 1, 200, (2, 2, 2))[3] == 100
    assert port.step_state((0, 0, 0), 0, 2, 200, (2, 2, 2))[3] == -0.1


def test_stateful_env_matches_pure_function():
    env = port.PortHanoi(3, 5)
    with pytest.raises(AssertionError):
        env.step(0)
    obs = env.reset()
    assert obs.dtype == np.float64 and obs.tolist() == [1, 0, 0] * 3
    out = [env.step(a) for a in (2, 2, 2, 2)]
    assert all(o[3] and not o[2] for o in out)
    o = env.step(2)  # 5th step: truncation coincides with an illegal move
    assert o[2] and o[3] and env.step_counter == 0 and not env.reset_check


@pytest.mark.parametrize("name", SEARCH_FIXTURES)
def test_search_injected_matches_reference(golden, name):
    """Tree arithmetic only (network outputs injected from the reference trace): visit counts,
    root value and the persistent min/max must equal the reference bit for bit."""
    g = golden(f"search_{name}.npz")
    S, K = int(g["S"]), int(g["K"])
    mm = port.MinMax()
    for k in range(K):
        prior = g["prior"][k] if bool(g["prior_is_f64"]) else g["prior"][k].astype(np.float32)
        tr = port.SearchTrace()
        search = port.PortSearch(float(g["discount"]), S, mm)
        visits, q, _ = search.run(prior, None, None, injected=(g["r"][k], g["p"][k], g["v"][k]), trace=tr)
        assert np.array_equal(visits, g["child_N"][k])
        assert q == g["root_q"][k]
        assert (mm.minimum, mm.maximum) == (g["mm_min"][k], g["mm_max"][k])
        for s in range(S):
            d = int(g["depth"][k, s])
            assert tr.actions_path[s] == g["path"][k, s, :d].tolist()
        pi = port.play_policy(visits, float(g["temperature"]))
        assert np.array_equal(pi, g["pi"][k])
        a = int(np.argmax(visits)) if bool(g["deterministic"]) else port.sample_action(pi, g["uniform"][k])
        assert a == g["action"][k]


@pytest.mark.parametrize("name", ["n3_s50_noise_t1", "n5_s100_noise_t05"])
def test_search_own_network_matches_reference(golden, name):
    """Whole run_mcts through the port's own float32 torch network.  Bit-exact only when the host
    BLAS rounds like the one the fixture was made with, so the network outputs are probed first."""
    import torch

    g = golden(f"search_{name}.npz")
    n, S = int(g["N"]), int(g["S"])
    net = port.PortNet(port.make_weights(n, int(g["weight_seed"])))
    h0, _, p0, v0 = net.initial_inference(torch.from_numpy(g["obs"][0]).to(torch.float32))
    if not (np.array_equal(h0, g["h0"][0]) and np.array_equal(p0, g["p0"][0])):
        pytest.skip("host float32 GEMV rounds differently from the fixture's host")
    mm = port.MinMax()
    for k in range(min(3, int(g["K"]))):
        a, pi, q, visits, _ = port.run_mcts_port(
            g["obs"][k], net, port.PortSearch(float(g["discount"]), S, mm), float(g["temperature"]),
            bool(g["deterministic"]), alpha=float(g["alpha"]), noise=g["noise"][k], u=g["uniform"][k])
        assert np.array_equal(visits, g["child_N"][k]) and q == g["root_q"][k] and a == g["action"][k]


def test_network_port_close_to_reference(golden):
    import torch

    g = golden("net_io.npz")
    for n in (3, 5, 10):
        net = port.PortNet(port.make_weights(n, int(g[f"n{n}_weight_seed"])))
        for i in range(0, 96, 8):
            a1 = torch.zeros(6)
            a1[int(g[f"n{n}_action"][i])] = 1.0
            h2, r, p, v = net.recurrent_inference(torch.from_numpy(g[f"n{n}_h_in"][i]), a1)
            np.testing.assert_allclose(h2, g[f"n{n}_h_out"][i], rtol=0, atol=2e-6)
            np.testing.assert_allclose(p, g[f"n{n}_p"][i], rtol=1e-5, atol=1e-7)
            assert abs(r - g[f"n{n}_r"][i]) <= 1e-5 * max(1.0, abs(g[f"n{n}_r"][i]))
            assert abs(v - g[f"n{n}_v"][i]) <= 1e-5 * max(1.0, abs(g[f"n{n}_v"][i]))


def test_ucb_table_and_policy_helpers():
    t = port.ucb_table(4)
    assert t[0] == 0.0 and t.dtype == np.float64
    assert np.array_equal(port.play_policy([1, 2, 1, 0, 0, 0], 0.0), np.array([1, 2, 1, 0, 0, 0]) / 4)
    assert np.array_equal(port.play_policy([1, 2, 1, 0, 0, 0], 0.5), np.array([1, 4, 1, 0, 0, 0]) / 6)
    with pytest.raises(ValueError):
        port.play_policy([1, 1, 1, 1, 1, 1], 1.5)
    assert port.sample_action(np.array([0.5, 0.5, 0, 0, 0, 0]), 0.5) == 1
    assert port.sample_action(np.array([0.5, 0.5, 0, 0, 0, 0]), 0.49) == 0


def test_n_step_returns():
    r = port.n_step_returns([0, 0, 100], [1.0, 2.0, 3.0], 2, 0.5)
    assert r == [0 + 0.5 * 0 + 0.25 * 3.0, 0 + 0.5 * 100 + 0.25 * 0, 100 + 0 + 0]


# ------------------------------------------------------------- C oracle (fast checker) pinning
@pytest.mark.parametrize("name", SEARCH_FIXTURES)
def test_c_oracle_search_matches_reference(golden, name):
    from oracle import cport

    g = golden(f"search_{name}.npz")
    S, K = int(g["S"]), int(g["K"])
    inf = float("inf")
    mm = np.stack([np.concatenate([[inf], g["mm_min"][:-1]]), np.concatenate([[-inf], g["mm_max"][:-1]])], 1).copy()
    visits, q, depth = cport.search_injected(
        g["prior"], bool(g["prior_is_f64"]), mm, np.ascontiguousarray(g["r"].T), np.ascontiguousarray(g["p"].transpose(1, 0, 2)),
        np.ascontiguousarray(g["v"].T), float(g["discount"]), port.ucb_table(S + 1), want_depth=True)
    assert np.array_equal(visits, g["child_N"]) and np.array_equal(q, g["root_q"])
    assert np.array_equal(mm[:, 0], g["mm_min"]) and np.array_equal(mm[:, 1], g["mm_max"])
    assert np.array_equal(depth.T, g["depth"])


@pytest.mark.parametrize("n", [3, 5, 7, 10])
def test_c_oracle_env_matches_reference_tables(golden, n):
    from oracle import cport

    g = golden("env_tables.npz")
    states = np.array([port.state_to_packed(port.index_to_state(i, n)) for i in range(3 ** n)], dtype=np.uint32)
    lut = np.zeros(1 << (2 * n), dtype=np.int64)
    lut[states] = np.arange(3 ** n)
    for vi in range(g[f"n{n}_obs_idx"].shape[0]):
        c0 = int(g["counter_before"][vi])
        words = (np.repeat(states, 6) | np.uint32(c0 << (2 * n))).astype(np.uint32)
        acts = np.tile(np.arange(6, dtype=np.uint8), 3 ** n)
        rew, flags, obs = cport.env_step(words, acts, n, int(g["max_steps"]))
        shape = (3 ** n, 6)
        assert np.array_equal(lut[obs].reshape(shape), g[f"n{n}_obs_idx"][vi])
        assert np.array_equal(lut[words & ((1 << (2 * n)) - 1)].reshape(shape), g[f"n{n}_stored_idx"][vi])
        assert np.array_equal((words >> (2 * n)).reshape(shape), g[f"n{n}_counter_after"][vi])
        assert np.array_equal((flags & 1).reshape(shape), g[f"n{n}_done"][vi])
        assert np.array_equal(((flags >> 1) & 1).reshape(shape), g[f"n{n}_illegal"][vi])
        code = np.where(rew == 100, 1, np.where(rew == 0, 0, 2))
        assert np.array_equal(code.reshape(shape), g[f"n{n}_reward_code"][vi])
    assert np.array_equal(cport.legal_mask(states, n), g[f"n{n}_legal"])
    assert np.array_equal(cport.solver(states, n), g[f"n{n}_solver"])


# ------------------------------------------------- episode post-processing (§8f row 1) pinning
def _episode(g, k):
    rw = [{0: 0, 1: 100, 2: -100 / 1000}[int(c)] for c in g[f"e{k}_reward_code"]]
    pis = [port.play_policy(v, 1.0) for v in g[f"e{k}_visits"]]
    return rw, [float(x) for x in g[f"e{k}_root_q"]], [int(a) for a in g[f"e{k}_action"]], pis


def test_episode_post_matches_reference(golden):
    """n-step returns (utils.py:28-72), priorities (Muzero.py:197-200) and organise_transitions
    (Muzero.py:276-323) of the port against outputs of the reference's own functions."""
    g = golden("episode_post.npz")
    n_step, discount, unroll = int(g["n_step"]), float(g["discount"]), int(g["unroll"])
    for k in range(int(g["n_episodes"])):
        rw, root_q, actions, pis = _episode(g, k)
        returns = port.n_step_returns(rw, root_q, n_step, discount)
        assert np.array_equal(np.array(returns, np.float64), g[f"e{k}_returns"])
        assert np.array_equal(port.priorities(returns, root_q), g[f"e{k}_priority"])
        assert np.array_equal(np.array(port.mc_returns(rw, discount)), g[f"e{k}_mc_returns"])
        assert np.array_equal(port.priorities(port.mc_returns(rw, discount), root_q), g[f"e{k}_mc_priority"])
        states = [np.zeros(15)] * len(rw)
        _, o_r, o_a, o_p, o_g = port.organise_transitions(states, rw, actions, pis, returns, unroll, 6, int(g[f"e{k}_absorbing"]))
        for name, arr in (("o_r", o_r), ("o_a", o_a), ("o_p", o_p), ("o_g", o_g)):
            assert arr.dtype == g[f"e{k}_{name}"].dtype and np.array_equal(arr, g[f"e{k}_{name}"]), (k, name)
