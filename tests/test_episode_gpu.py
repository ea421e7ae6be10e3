"""GPU parity: episode post-processing kernels (n-step returns, priorities, organise_transitions into the
replay ring) against outputs of the reference's own functions (tests/golden/episode_post.npz, generated
from the unmodified utils.compute_n_step_returns / Muzero.organise_transitions) — bit-exact."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu
FLAG = {0: 0, 1: 4 | 1, 2: 2}  # reward code -> HMZ_FLAG_* of the move (goal implies done)


def _fill_store(g, n_disks=5, t_max=200, order=None):
    from muzero_hanoi_b200.replay import EpisodeStore

    K = int(g["n_episodes"])
    order = list(range(K)) if order is None else order
    st = EpisodeStore(len(order), t_max, n_disks)
    rng = np.random.default_rng(5)
    words = np.zeros((t_max, len(order)), np.int32)
    ep_len = np.zeros(len(order), np.int32)
    for col, k in enumerate(order):
        codes = g[f"e{k}_reward_code"]
        T = len(codes)
        ep_len[col] = T
        words[:T, col] = [port.state_to_packed(port.index_to_state(int(i), n_disks)) for i in rng.integers(0, 242, T)]
        st.flags[:T, col] = torch.tensor([FLAG[int(c)] for c in codes], dtype=torch.uint8)
        st.action[:T, col] = torch.from_numpy(g[f"e{k}_action"].astype(np.uint8))
        st.visits[:T, col] = torch.from_numpy(g[f"e{k}_visits"].astype(np.int16))
        st.root_q[:T, col] = torch.from_numpy(g[f"e{k}_root_q"])
    st.state.copy_(torch.from_numpy(words))
    st.ep_len.copy_(torch.from_numpy(ep_len))
    return st, words, ep_len


def test_returns_and_priorities_bit_exact(golden):
    g = golden("episode_post.npz")
    st, _, ep_len = _fill_store(g)
    ret, prio = st.post_process(int(g["n_step"]), float(g["discount"]))
    torch.cuda.synchronize()
    ret, prio = ret.cpu().numpy(), prio.cpu().numpy()
    for k in range(int(g["n_episodes"])):
        T = ep_len[k]
        assert np.array_equal(ret[:T, k], g[f"e{k}_returns"]), k       # float64, bit for bit
        assert np.array_equal(prio[:T, k], g[f"e{k}_priority"]), k     # float32, bit for bit


def test_mc_returns_and_priorities_bit_exact(golden):
    """TD_return=False branch: compute_MCreturns (utils.py:75-86) incl. NumPy's own pow for the discounts."""
    g = golden("episode_post.npz")
    st, _, ep_len = _fill_store(g)
    ret, prio = st.post_process_mc(float(g["discount"]))
    torch.cuda.synchronize()
    ret, prio = ret.cpu().numpy(), prio.cpu().numpy()
    for k in range(int(g["n_episodes"])):
        T = ep_len[k]
        assert np.array_equal(ret[:T, k], g[f"e{k}_mc_returns"]), k
        assert np.array_equal(prio[:T, k], g[f"e{k}_mc_priority"]), k


@pytest.mark.parametrize("capacity", [1000, 300])
def test_unroll_into_replay_ring_matches_organise_transitions(golden, capacity):
    """Rows land in episode order from ptr with buffer.py's wrap-around; content equals the reference's arrays."""
    from muzero_hanoi_b200.replay import ReplayRing

    g = golden("episode_post.npz")
    order = [5, 0, 3, 7, 1] if capacity == 1000 else [4, 6, 2]  # lengths 24,1,7,200,2 / 11,60,5
    st, words, ep_len = _fill_store(g, order=order)
    st.post_process(int(g["n_step"]), float(g["discount"]))
    ring = ReplayRing(capacity, int(g["unroll"]), 15, 6)
    ring.ptr = capacity - 40  # force a wrap
    start = ring.ptr
    absorbing = np.array([int(g[f"e{k}_absorbing"]) for k in order], np.uint8)
    n = ring.add_episodes(st, temperature=1.0, only_solved=False, absorbing_action=absorbing)
    torch.cuda.synchronize()
    assert n == int(ep_len.sum()) and ring.ptr == (start + n) % capacity and ring.is_full
    row = start
    for col, k in enumerate(order):
        T = ep_len[col]
        rows = (row + np.arange(T)) % capacity
        assert np.array_equal(ring.rwds.cpu().numpy()[rows], g[f"e{k}_o_r"])
        assert np.array_equal(ring.actions.cpu().numpy()[rows], g[f"e{k}_o_a"])
        assert np.array_equal(ring.pi_probs.cpu().numpy()[rows], g[f"e{k}_o_p"])
        assert np.array_equal(ring.mc_returns.cpu().numpy()[rows], g[f"e{k}_o_g"])
        assert np.array_equal(ring.priorities.cpu().numpy()[rows], g[f"e{k}_priority"])
        onehot = np.stack([port.one_hot(port.packed_to_state(int(w), 5)) for w in words[:T, col]]).astype(np.float32)
        assert np.array_equal(ring.states.cpu().numpy()[rows], onehot)
        row += T


@pytest.mark.parametrize("capacity,scalar", [(100, False), (37, False), (100, True)])
def test_more_rows_than_the_ring_holds_equals_sequential_adds(golden, capacity, scalar, monkeypatch):
    """All eight fixture episodes (310 rows) finish on one move and that fit go into a ring of 100 / 37 rows (110 / 50 rows arrive).  The reference adds
    episodes one at a time (Muzero.py:98-101 -> buffer.py:47-83), later rows overwriting earlier ones; the device-side
    add of the whole batch must leave exactly the ring those sequential adds leave (ptr and is_full included)."""
    from muzero_hanoi_b200.replay import ReplayRing

    if scalar:
        monkeypatch.setenv("HMZ_UNROLL_SCALAR", "1")
    g = golden("episode_post.npz")
    order = [k for k in range(int(g["n_episodes"])) if len(g[f"e{k}_returns"]) <= capacity]
    st, words, ep_len = _fill_store(g, order=order)
    st.post_process(int(g["n_step"]), float(g["discount"]))
    absorbing = np.array([int(g[f"e{k}_absorbing"]) for k in order], np.uint8)
    dev, seq = ReplayRing(capacity, int(g["unroll"]), 15, 6), ReplayRing(capacity, int(g["unroll"]), 15, 6)
    dev.ptr = seq.ptr = 11
    n = dev.add_episodes(st, temperature=1.0, only_solved=False, absorbing_action=absorbing)
    assert n == int(ep_len.sum()) and n > capacity
    for col, k in enumerate(order):  # the reference's order of operations, through the reference-shaped host add
        T = ep_len[col]
        onehot = np.stack([port.one_hot(port.packed_to_state(int(w), 5)) for w in words[:T, col]]).astype(np.float32)
        seq.add(onehot, g[f"e{k}_o_r"], g[f"e{k}_o_a"], g[f"e{k}_o_p"], g[f"e{k}_o_g"], g[f"e{k}_priority"])
    torch.cuda.synchronize()
    assert dev.ptr == seq.ptr and dev.is_full and seq.is_full
    for name in ("states", "rwds", "actions", "pi_probs", "mc_returns", "priorities"):
        assert torch.equal(getattr(dev, name), getattr(seq, name)), name


def test_only_solved_filter_and_row_bases(golden):
    """training_loop keeps an episode only if returns[-1, 0] > 0 (Muzero.py:98)."""
    from muzero_hanoi_b200.replay import ReplayRing

    g = golden("episode_post.npz")
    st, _, ep_len = _fill_store(g)
    ret, _ = st.post_process(int(g["n_step"]), float(g["discount"]))
    ring = ReplayRing(2000, int(g["unroll"]), 15, 6)
    n = ring.add_episodes(st, only_solved=True)
    torch.cuda.synchronize()
    keep = [k for k in range(len(ep_len)) if g[f"e{k}_returns"][-1] > 0]
    assert n == sum(int(ep_len[k]) for k in keep) and len(ring) == n
    base = st.row_base.cpu().numpy()
    expect, at = [], 0
    for k in range(len(ep_len)):
        expect.append(at if k in keep else -1)
        at += int(ep_len[k]) if k in keep else 0
    assert base.tolist() == expect


def test_selfplay_episode_store_agrees_with_oracle_returns():
    """End to end: SelfPlay records moves into the episode store; finished episodes' returns equal the
    oracle's compute_n_step_returns on the recorded rewards and root values, and Buffer-style sampling works."""
    from muzero_hanoi_b200 import _lib
    from muzero_hanoi_b200.engine import PackedWeights, SelfPlay
    from muzero_hanoi_b200.replay import ReplayRing

    n, B, S, max_steps = 3, 256, 16, 12
    w = PackedWeights(port.make_weights(n, 2), n, _lib.MODE_FP32)
    sp = SelfPlay(n, max_steps, B, S, w, seed=3, episodes=True)
    ring = ReplayRing(20000, 5, 3 * n, 6)
    seen = 0
    for _ in range(2 * max_steps):
        sp.move()
        st = sp.episodes
        ret, prio = st.post_process(10, 0.8)
        torch.cuda.synchronize()
        ep_len = st.ep_len.cpu().numpy()
        flags, rq, rets = st.flags.cpu().numpy(), st.root_q.cpu().numpy(), ret.cpu().numpy()
        for gidx in np.nonzero(ep_len)[0][:8]:
            T = ep_len[gidx]
            rw = [100 if f & 4 else (-100 / 1000 if f & 2 else 0) for f in flags[:T, gidx]]
            assert rets[:T, gidx].tolist() == port.n_step_returns(rw, [float(x) for x in rq[:T, gidx]], 10, 0.8)
            assert flags[T - 1, gidx] & 1 and not (flags[:T - 1, gidx] & 1).any()
            seen += 1
        ring.add_episodes(st, only_solved=False)
    assert seen > 0 and len(ring) > 0
    np.random.seed(0)
    states, rwds, actions, pi, g_, indx, wts = ring.priority_sample(64)
    assert states.shape == (64, 3 * n) and pi.shape == (64, 5, 6) and wts.shape == (64,) and states.is_cuda
    assert torch.allclose(pi.sum(-1), torch.ones_like(pi.sum(-1)), atol=1e-6)
    ring.update_priorities(indx, np.full(64, 0.5, np.float32))
    assert np.allclose(ring.priorities[torch.as_tensor(indx, device="cuda")].cpu().numpy(), 0.5)
