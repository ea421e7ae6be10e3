"""GPU parity: libhmz env kernels (through the C ABI) vs the golden transition tables generated
from the unmodified reference, vs the oracle port on seeded inputs, and — at BASELINE.json's
full config-4 size (N=10, 2^24 envs) — through size-independent properties.  Bit-exact."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu
REWARD_F32 = {0: np.float32(0.0), 1: np.float32(100.0), 2: np.float32(-0.1)}


def _table_words(n, counter):
    idx = np.repeat(np.arange(3 ** n, dtype=np.int64), 6)
    act = np.tile(np.arange(6, dtype=np.uint8), 3 ** n)
    states = np.array([port.state_to_packed(port.index_to_state(i, n)) for i in range(3 ** n)], dtype=np.uint32)
    words = (np.repeat(states, 6) | (np.uint32(counter) << np.uint32(2 * n))).astype(np.uint32)
    return idx, act, words, states


@pytest.mark.parametrize("n", [3, 4, 5, 7, 10])
def test_step_exhaustive_vs_reference_tables(golden, n):
    from muzero_hanoi_b200.engine import VecHanoi

    g = golden("env_tables.npz")
    max_steps = int(g["max_steps"])
    variants = range(g[f"n{n}_obs_idx"].shape[0])
    for vi in variants:
        c0 = int(g["counter_before"][vi])
        idx, act, words, states = _table_words(n, c0)
        env = VecHanoi(n, max_steps, len(words), auto_reset=False)
        env.words.copy_(torch.from_numpy(words.view(np.int32)))
        obs_w, rew, flags = env.step(torch.from_numpy(act).cuda())
        st_after, ctr_after = env.states()
        obs_w = obs_w.cpu().numpy().view(np.uint32)
        flags = flags.cpu().numpy()
        rew = rew.cpu().numpy()
        lut = np.zeros(1 << (2 * n), dtype=np.int64)
        lut[states] = np.arange(3 ** n)
        shape = (3 ** n, 6)
        assert np.array_equal(lut[obs_w].reshape(shape), g[f"n{n}_obs_idx"][vi])
        assert np.array_equal(lut[st_after].reshape(shape), g[f"n{n}_stored_idx"][vi])
        assert np.array_equal(ctr_after.reshape(shape), g[f"n{n}_counter_after"][vi])
        assert np.array_equal((flags & 1).reshape(shape), g[f"n{n}_done"][vi])
        assert np.array_equal(((flags >> 1) & 1).reshape(shape), g[f"n{n}_illegal"][vi])
        want_r = np.vectorize(REWARD_F32.get)(g[f"n{n}_reward_code"][vi]).astype(np.float32)
        assert np.array_equal(rew.reshape(shape), want_r)
        # done <=> reset_check cleared in the reference
        assert np.array_equal((flags & 1).reshape(shape), 1 - g[f"n{n}_reset_check"][vi])
        # goal flag <=> reward 100; trunc flag <=> counter hit max_steps
        assert np.array_equal(((flags >> 2) & 1).reshape(shape), g[f"n{n}_reward_code"][vi] == 1)


@pytest.mark.parametrize("n", [3, 5, 10])
def test_legal_mask_solver_onehot_index(golden, n):
    from muzero_hanoi_b200.engine import VecHanoi

    g = golden("env_tables.npz")
    env = VecHanoi(n, 200, 3 ** n)
    env.set_state_indices(np.arange(3 ** n))
    states, ctr = env.states()
    want = np.array([port.state_to_packed(port.index_to_state(i, n)) for i in range(3 ** n)], dtype=np.uint32)
    assert np.array_equal(states, want) and not ctr.any()
    assert np.array_equal(env.state_indices().cpu().numpy(), np.arange(3 ** n))
    assert np.array_equal(env.legal_mask().cpu().numpy(), g[f"n{n}_legal"])
    assert np.array_equal(env.solver_distance().cpu().numpy(), g[f"n{n}_solver"].astype(np.int32))
    oh = env.onehot().cpu().numpy()
    pick = np.random.default_rng(0).integers(0, 3 ** n, 200)
    for i in pick:
        assert np.array_equal(oh[i].astype(np.float64), port.one_hot(port.index_to_state(int(i), n)))


@pytest.mark.parametrize("b", [0, 1, 3, 5, 1027])
def test_ragged_and_empty_batches(b):
    from muzero_hanoi_b200.engine import VecHanoi

    n, max_steps = 4, 9
    env = VecHanoi(n, max_steps, b, auto_reset=True)
    env.reset()
    rng = np.random.default_rng(b)
    oracle = [((0,) * n, 0) for _ in range(b)]
    for t in range(25):
        acts = rng.integers(0, 6, b).astype(np.uint8)
        _, rew, flags = env.step(torch.from_numpy(acts).cuda(), want_obs=False)
        st, ctr = env.states()
        rew, flags = rew.cpu().numpy(), flags.cpu().numpy()
        for i in range(b):
            s, c = oracle[i]
            moved, stored, c2, r, d, ill, _ = port.step_state(s, c, int(acts[i]), max_steps, (2,) * n)
            if d:
                stored, c2 = (0,) * n, 0  # auto-reset to init_state_idx=0
            oracle[i] = (stored, c2)
            assert port.state_to_packed(stored) == st[i] and c2 == ctr[i]
            assert np.float32(r) == rew[i] and bool(flags[i] & 1) == d and bool(flags[i] & 2) == ill


def test_unaligned_views_take_the_scalar_path():
    from muzero_hanoi_b200 import _lib
    from muzero_hanoi_b200.engine import VecHanoi

    lib = _lib.load()
    n, b = 3, 64
    env = VecHanoi(n, 200, b + 1)
    env.reset()
    words = env.words[1:]  # 4-byte aligned only
    acts = torch.full((b + 1,), 1, dtype=torch.uint8, device="cuda")[1:]
    rew = torch.zeros(b + 1, device="cuda")[1:]
    flags = torch.zeros(b + 1, dtype=torch.uint8, device="cuda")[1:]
    _lib.check(lib.hmz_env_step(_lib.ptr(words), _lib.ptr(acts), _lib.ptr(rew), _lib.ptr(flags), None, b, n, 200, 2,
                                0, 0, _lib.current_stream()))
    torch.cuda.synchronize()
    assert (words.cpu().numpy().view(np.uint32) == (port.state_to_packed((2, 0, 0)) | (1 << 6))).all()
    assert int(env.words[0].item()) == 0


def test_dropin_env_matches_port_random_walks():
    from muzero_hanoi_b200.env.hanoi import TowersOfHanoi
    from muzero_hanoi_b200.env.hanoi_utils import hanoi_solver

    rng = np.random.default_rng(3)
    for n, max_steps in ((3, 12), (5, 40)):
        env, ref = TowersOfHanoi(n, max_steps), port.PortHanoi(n, max_steps)
        with pytest.raises(AssertionError):
            env.step(0)
        for ep in range(6):
            o1, o2 = env.reset(), ref.reset()
            assert o1.dtype == np.float64 and np.array_equal(o1, o2)
            done = False
            while not done:
                a = int(rng.integers(0, 6))
                assert env._move_allowed(env.moves[a]) == port.move_allowed(ref.c_state, port.MOVES[a])
                x, y = env.step(a), ref.step(a)
                assert np.array_equal(x[0], y[0]) and x[0].dtype == np.float64
                assert x[1] == y[1] and type(x[1]) is type(y[1]) and x[2:] == y[2:]
                assert env.c_state == ref.c_state and env.step_counter == ref.step_counter
                assert env.reset_check == ref.reset_check and env.current_state() == ref.current_state()
                done = x[2]
            with pytest.raises(AssertionError):
                env.step(0)
    assert [hanoi_solver(s) for s in ((2, 2, 0), (0, 0, 2), (1, 2, 2), (0, 0, 0))] == [7, 3, 1, 7]
    env = TowersOfHanoi(3, 200)
    env.reset()
    assert env.states.index((2, 2, 0)) == 24 and env.states[1] == (0, 0, 1) and len(env.states) == 27
    assert env._get_moved_state((0, 1)) == (1, 0, 0)
    with pytest.raises(UnboundLocalError):
        env._get_moved_state((1, 0))
    with pytest.raises(IndexError):
        env.step(6)


def test_random_reset_is_uniform_over_non_goal_states():
    from muzero_hanoi_b200.engine import VecHanoi

    env = VecHanoi(3, 200, 1 << 18)
    env.random_reset(seed=5)
    idx = env.state_indices().cpu().numpy()
    counts = np.bincount(idx, minlength=27)
    assert counts[26] == 0 and counts[:26].min() > 0.9 * (1 << 18) / 26 and counts[:26].max() < 1.1 * (1 << 18) / 26


def test_step_random_matches_port_given_same_actions():
    """On-device random legal moves: every chosen action is legal for the state it was drawn in,
    and replaying the chosen actions through the oracle reproduces the device states."""
    from muzero_hanoi_b200.engine import VecHanoi

    n, b, max_steps = 5, 256, 50
    env = VecHanoi(n, max_steps, b)
    env.reset()
    oracle = [((0,) * n, 0)] * b
    for t in range(80):
        acts, rew, flags = env.step_random(seed=11, step_index=t)
        acts, flags = acts.cpu().numpy(), flags.cpu().numpy()
        st, ctr = env.states()
        assert not (flags & 2).any()
        for i in range(b):
            s, c = oracle[i]
            assert port.move_allowed(s, port.MOVES[acts[i]])
            _, stored, c2, r, d, ill, _ = port.step_state(s, c, int(acts[i]), max_steps, (2,) * n)
            if d:
                stored, c2 = (0,) * n, 0
            oracle[i] = (stored, c2)
            assert port.state_to_packed(stored) == st[i] and c2 == ctr[i]
    # all three legal moves get used from the start state's successors
    assert len(set(acts.tolist())) >= 3


def test_step_random_is_uniform_over_the_legal_moves():
    """The on-device move choice (two moves of the smallest disk + at most one move between the other two pegs) must be
    uniform over exactly the legal set: 600,000 envs in one state, chi-square-style 5-sigma bounds per action."""
    from muzero_hanoi_b200.engine import VecHanoi

    n, b = 4, 600_000
    for state in ((0, 0, 0, 0), (1, 0, 2, 2), (2, 1, 1, 0), (0, 2, 2, 2), (1, 1, 1, 1), (2, 0, 1, 2)):
        legal = [a for a in range(6) if port.move_allowed(state, port.MOVES[a])]
        env = VecHanoi(n, 200, b)
        env.set_state_indices(np.full(b, port.state_to_index(state), dtype=np.int32))
        acts, _, flags = env.step_random(seed=3, step_index=17)
        counts = np.bincount(acts.cpu().numpy(), minlength=6)
        assert not (flags.cpu().numpy() & 2).any()
        assert sorted(np.nonzero(counts)[0].tolist()) == legal, (state, counts)
        p = 1.0 / len(legal)
        assert (np.abs(counts[legal] - b * p) < 5 * np.sqrt(b * p * (1 - p))).all(), (state, counts)


def test_fused_rollout_equals_stepwise_random():
    from muzero_hanoi_b200.engine import VecHanoi

    n, b, k = 10, 1 << 16, 37
    a, c = VecHanoi(n, 200, b), VecHanoi(n, 200, b)
    a.reset(), c.reset()
    goals = truncs = 0
    for t in range(k):
        _, _, flags = a.step_random(seed=7, step_index=t)
        goals += int((flags & 4).ne(0).sum())
        truncs += int((flags & 8).ne(0).sum())
    counters = c.rollout_random(k, seed=7, step_index=0).cpu().numpy()
    assert torch.equal(a.words, c.words)
    assert counters[0] == b * k and counters[1] == goals and counters[2] == truncs


def test_full_size_config4_properties():
    """BASELINE.json config 4 at full size (N=10, 2^24 envs): properties that do not need the
    oracle — never an illegal move, step accounting exact, episodes end only by truncation or
    goal, every state stays a valid packing, and solved episodes need >= 2^N - 1 moves."""
    from muzero_hanoi_b200.engine import VecHanoi

    n, b, max_steps = 10, 1 << 24, 200
    env = VecHanoi(n, max_steps, b)
    env.reset()
    total_trunc = 0
    for t in range(0, 400, 100):
        c = env.rollout_random(100, seed=1, step_index=t).cpu().numpy()
    assert c[0] == b * 400
    assert c[1] == 0  # 1023 moves are needed from the start state; max_steps = 200 truncates first
    assert c[2] == b * 2  # every env truncates exactly at steps 200 and 400
    st, ctr = env.states()
    assert (ctr == 0).all() and (st == 0).all()  # all just auto-reset
    env.rollout_random(63, seed=1, step_index=400)
    st, ctr = env.states()
    assert (ctr == 63).all()
    lo, hi = st & 0x55555555, (st >> 1) & 0x55555555
    assert not (lo & hi).any()  # no disk on "peg 3"
    _, _, flags = env.step_random(seed=1, step_index=463)
    assert not (flags & 2).any() and not (flags & 1).any()


def test_short_puzzle_reaches_goal_with_random_moves():
    from muzero_hanoi_b200.engine import VecHanoi

    n, b = 3, 1 << 16
    env = VecHanoi(n, 200, b)
    env.reset()
    c = env.rollout_random(200, seed=2).cpu().numpy()
    assert c[0] == b * 200 and c[1] > b  # random walks solve N=3 (7 moves optimal) often
    # a solved episode takes at least 7 steps, so at most 200/7 goals per env
    assert c[1] <= b * (200 // 7)
