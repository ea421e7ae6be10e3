"""GPU parity of the whole search (MCTS.run_mcts): BASELINE.json config 2 semantics —
N=3, 4,096 parallel searches x 50 simulations — plus the drop-in MCTS class.

Bit-exact gate: the device search (own float32 network) records the network outputs of every
simulation; the C oracle (pinned to the reference, tests/test_oracle_golden.py) replays the tree
arithmetic with those outputs injected and must reproduce visit counts, root values, leaf depths
and the persistent min/max bit for bit.  The network outputs themselves are gated separately
(tests/test_net_gpu.py, <= 1e-5)."""
import numpy as np
import pytest
import torch

from oracle import cport, port

pytestmark = pytest.mark.gpu
# hmz_search_t.schedule values exercised at full size: serial launches, the automatic choice (4 stream groups at this
# size), 7 ragged groups, and the persistent role-specialised kernel
SCHEDULES_FULL = (1, 0, 7, 64, 128)


def _split_search(mcts, weights, words, noise, record=True):
    """BatchedMCTS.run_mcts unrolled into its phases so that per-simulation outputs can be recorded."""
    import ctypes as C

    from muzero_hanoi_b200 import _lib

    st, B, S = mcts.store, mcts.B, mcts.n_simulations
    dev = mcts.device
    p0, v0 = torch.empty(B, 6, device=dev), torch.empty(B, device=dev)
    weights.initial(B, words=words, latents_out=st.latents, out_rows_per_item=st.n_records, latent_dtype=0, p0=p0, v0=v0)
    nz = None if noise is None else torch.from_numpy(noise).cuda()
    st.desc.root_prior_is_f64 = int(noise is not None)
    _lib.check(mcts.lib.hmz_search_begin_p0(C.byref(st.desc), _lib.ptr(p0), _lib.ptr(nz), 0.25, _lib.current_stream()))
    r, v, p = torch.empty(S, B, device=dev), torch.empty(S, B, device=dev), torch.empty(S, B, 6, device=dev)
    depth = torch.zeros(S, B, dtype=torch.int16, device=dev)
    for s in range(S):
        mcts.select(s)
        depth[s].copy_(mcts.leaf_depth)
        weights.recurrent(B, latents_in=st.latents, in_rows_per_item=st.n_records, in_row=mcts.leaf_parent,
                          actions=mcts.leaf_action, latents_out=st.latents, out_rows_per_item=st.n_records,
                          out_row=s + 1, latent_dtype=0, r=r[s], p=p[s], v=v[s])
        mcts.expand_backup(s, r[s], p[s], v[s])
    return p0, r, p, v, depth


@pytest.mark.parametrize("n,B,S,use_noise", [(3, 4096, 50, True), (4, 1024, 200, False), (5, 2048, 100, True)])
def test_config2_visit_counts_bit_exact_vs_oracle(n, B, S, use_noise):
    from muzero_hanoi_b200.engine import BatchedMCTS, PackedWeights, VecHanoi

    weights = PackedWeights(port.make_weights(n, 0), n)
    env = VecHanoi(n, 200, B)
    non_goal = np.array([i for i in range(3 ** n) if i != 3 ** n - 1])
    env.set_state_indices(non_goal[np.arange(B) % len(non_goal)].astype(np.int32))
    rng = np.random.default_rng(0)
    noise = rng.dirichlet(np.full(6, 0.25), size=B) if use_noise else None
    mcts = BatchedMCTS(0.8, 0.25 if use_noise else 0.0, S, B)
    for move in range(2):  # second move: MinMaxStats carried over, like a reused MCTS object
        mm_before = mcts.store.minmax.cpu().numpy().copy()
        p0, r, p, v, depth = _split_search(mcts, weights, env.words, noise)
        act, pi, q, visits = mcts.root_policy(1.0, False, uniforms=rng.random(B))
        torch.cuda.synchronize()
        prior = mcts.store.root_prior.cpu().numpy()
        if use_noise:
            assert np.array_equal(prior, port.mix_dirichlet(p0.cpu().numpy(), noise))
        else:
            assert np.array_equal(prior, p0.cpu().numpy().astype(np.float64))
        mm = mm_before.copy()
        o_visits, o_q, o_depth = cport.search_injected(prior, use_noise, mm, r.cpu().numpy(), p.cpu().numpy(),
                                                       v.cpu().numpy(), 0.8, port.ucb_table(S + 1), want_depth=True)
        assert np.array_equal(visits.cpu().numpy(), o_visits)
        assert np.array_equal(q.cpu().numpy(), o_q)
        assert np.array_equal(depth.cpu().numpy().astype(np.uint16), o_depth)
        assert np.array_equal(mcts.store.minmax.cpu().numpy(), mm)
        assert (visits.sum(1) == S).all()
        env.step(act.to(torch.uint8), want_obs=False)


def test_fused_run_equals_split_phases():
    """hmz_search_run (one call) == select / recurrent / expand_backup issued one by one."""
    from muzero_hanoi_b200.engine import BatchedMCTS, PackedWeights, VecHanoi

    n, B, S = 3, 777, 30
    weights = PackedWeights(port.make_weights(n, 5), n)
    env = VecHanoi(n, 200, B)
    env.random_reset(seed=3)
    noise = np.random.default_rng(1).dirichlet(np.full(6, 0.25), size=B)
    a, b = BatchedMCTS(0.8, 0.25, S, B), BatchedMCTS(0.8, 0.25, S, B)
    _split_search(a, weights, env.words, noise)
    u = np.random.default_rng(2).random(B)
    ra = [t.clone() for t in a.root_policy(0.5, False, uniforms=u)]
    rb = b.run_mcts(weights, words=env.words, temperature=0.5, deterministic=False, noise=noise, uniforms=u)
    for x, y in zip(ra, rb):
        assert torch.equal(x, y)
    assert torch.equal(a.store.nodes, b.store.nodes) and torch.equal(a.store.minmax, b.store.minmax)


@pytest.mark.parametrize("name", ["n3_s50_noise_t1", "n3_s25_nonoise_t0", "n5_s100_noise_t05"])
def test_dropin_mcts_close_to_reference_episode(golden, name):
    """Drop-in MCTS + MuZeroNet + TowersOfHanoi replaying a reference episode with the reference's
    noise / uniform draws.  The device network is float32-close (not bit-identical) to torch's
    CPU kernels, so visit counts are compared statistically: root policy within 1e-5 and a mean
    absolute visit difference under 2% of the simulation budget; exact agreement is reported."""
    from muzero_hanoi_b200.MCTS.mcts import MCTS
    from muzero_hanoi_b200.networks import MuZeroNet

    g = golden(f"search_{name}.npz")
    n, S, K = int(g["N"]), int(g["S"]), int(g["K"])
    net = MuZeroNet(3 * n, 6, 0.002, "cpu", TD_return=True)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in port.make_weights(n, int(g["weight_seed"])).items()})
    mcts = MCTS(float(g["discount"]), float(g["alpha"]), S, 1, "cpu")

    class _Feed:  # supplies the reference's recorded draws through the np.random calls run_mcts makes
        def __init__(self):
            self.k = 0

        def dirichlet(self, alphas):
            return g["noise"][self.k]

        def random_sample(self):
            return g["uniform"][self.k]

    feed = _Feed()
    import muzero_hanoi_b200.MCTS.mcts as mod

    real = mod.np.random
    mod.np.random = feed
    try:
        diffs, exact = [], 0
        for k in range(K):
            feed.k = k
            action, pi, q = mcts.run_mcts(g["obs"][k], net, float(g["temperature"]), bool(g["deterministic"]))
            assert isinstance(action, int) and pi.dtype == np.float64 and pi.shape == (6,) and isinstance(q, float)
            visits = mcts._engine.visits.cpu().numpy()[0]
            assert visits.sum() == S
            assert np.abs(mcts._engine.p0.cpu().numpy()[0] - g["p0"][k]).max() <= 1e-5
            diffs.append(np.abs(visits - g["child_N"][k]).sum() / 2)
            exact += int(np.array_equal(visits, g["child_N"][k]))
            # keep the persistent MinMaxStats on the reference trajectory for the next move
            mcts.min_max_stats.minimum, mcts.min_max_stats.maximum = float(g["mm_min"][k]), float(g["mm_max"][k])
            la = mcts.return_latent_actions()
            assert 1 <= len(la) <= S and all(t.dtype == torch.long and t.shape == (1,) for t in la)
    finally:
        mod.np.random = real
    print(f"{name}: {exact}/{K} searches with identical visit counts, mean moved visits {np.mean(diffs):.2f} of {S}")
    assert np.mean(diffs) <= 0.02 * S + 0.5
    with pytest.raises(ValueError):
        mcts.run_mcts(g["obs"][0], net, 1.5, True)


@pytest.mark.parametrize("mode", [0, 2])  # parity mode on the FFMA kernel / on the tensor cores (HMZ_MODE_FP32X3)
def test_config5_lesion_mode_bit_exact_vs_oracle(mode):
    """BASELINE.json config 5 semantics (acting_ablations.py lesion mode): N=4, random non-goal starts,
    policy and value heads re-initialised U(-1/8, 1/8) (networks.py:201-205), no root noise, S=200,
    temperature 0 with sampling (deterministic=False) — 16,384 searches replayed by the C oracle."""
    from muzero_hanoi_b200.engine import BatchedMCTS, PackedWeights, VecHanoi

    n, B, S = 4, 16384, 200
    sd = port.lesion_weights(port.make_weights(n, 3), ("policy_net", "value_net"), seed=103)
    weights = PackedWeights(sd, n, mode)
    env = VecHanoi(n, 200, B)
    env.random_reset(seed=5)
    mcts = BatchedMCTS(0.8, 0.0, S, B)
    p0, r, p, v, depth = _split_search(mcts, weights, env.words, None)
    u = np.random.default_rng(1).random(B)
    act, pi, q, visits = mcts.root_policy(0.0, False, uniforms=u)
    torch.cuda.synchronize()
    mm = np.tile(np.array([[np.inf, -np.inf]]), (B, 1))
    o_visits, o_q, o_depth = cport.search_injected(p0.cpu().numpy().astype(np.float64), False, mm, r.cpu().numpy(),
                                                   p.cpu().numpy(), v.cpu().numpy(), 0.8, port.ucb_table(S + 1), want_depth=True)
    assert np.array_equal(visits.cpu().numpy(), o_visits) and np.array_equal(q.cpu().numpy(), o_q)
    assert np.array_equal(depth.cpu().numpy().astype(np.uint16), o_depth)
    assert int(depth.max()) > 8  # lesioned searches go deep: the parent-link backup path is exercised too
    pi_o = np.stack([port.play_policy(o_visits[i], 0.0) for i in range(256)])
    assert np.array_equal(pi.cpu().numpy()[:256], pi_o)
    assert [port.sample_action(pi_o[i], u[i]) for i in range(256)] == act.cpu().numpy()[:256].tolist()


def test_node_view_matches_tree_store(golden):
    """MCTS/node.py drop-in: Node views over the device tree expose the reference's fields / methods."""
    from muzero_hanoi_b200.MCTS.mcts import MCTS
    from muzero_hanoi_b200.MCTS.node import Node
    from muzero_hanoi_b200.networks import MuZeroNet

    g = golden("search_n3_s25_nonoise_t0.npz")
    net = MuZeroNet(9, 6, 0.002, "cpu", TD_return=True)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in port.make_weights(3, 1).items()})
    mcts = MCTS(0.8, 0.0, 25, 1, "cpu")
    mcts.run_mcts(g["obs"][0], net, 0.0, True)
    root = mcts.root_node()
    assert isinstance(root, Node) and root.is_expanded and not root.has_parent and root.prior == 0.0
    assert root.N == 25 and np.array_equal(root.child_N, g["child_N"][0]) and root.child_N.dtype == np.int32
    assert root.Q == g["root_q"][0] and root.rwd == 0.0 and root.h_state.shape == (64,)
    kids = root.children
    assert len(kids) == 6 and all(k.has_parent and k.parent is root and k.move == a for a, k in enumerate(kids))
    assert np.array_equal(np.array([k.prior for k in kids], np.float32), mcts._engine.p0.cpu().numpy()[0])
    assert np.abs(np.array([k.prior for k in kids], np.float32) - g["p0"][0]).max() <= 1e-5
    q, u = root.child_Q(mcts, mcts.min_max_stats), root.child_U(mcts)
    assert q.dtype == np.float32 and u.dtype == np.float32 and q.shape == (6,)
    # the reference formulas (node.py:90-123) on the same statistics
    mm = mcts.min_max_stats
    for a, k in enumerate(kids):
        want_q = np.float32(mm.normalize(k.rwd + 0.8 * k.Q)) if k.N > 0 else np.float32(0)
        w = (np.log((root.N + 19652 + 1) / 19652) + 1.25) * np.sqrt(root.N) / (k.N + 1)
        assert q[a] == want_q and u[a] == np.float32(np.float32(k.prior) * np.float32(w))
    best = root.best_child(mcts, mm)
    assert best.move == int(np.argmax(q + u))
    deep = best
    while deep.is_expanded and deep.children:
        nxt = deep.best_child(mcts, mm)
        if not nxt.is_expanded:
            break
        deep = nxt
    assert deep.N >= 1 and sum(c.N for c in deep.children) <= deep.N
    with pytest.raises(RuntimeError, match="already been expanded"):
        root.expand(g["p0"][0], None, 0.0)
    leaf = next(c for c in deep.children if not c.is_expanded)
    with pytest.raises(ValueError, match="Expand leaf node first"):
        leaf.best_child(mcts, mm)
    assert leaf.N == 0 and leaf.Q == 0.0 and leaf.h_state is None and leaf.children == []


@pytest.mark.parametrize("n,B,S,schedule", [(3, 4096, 50, 0), (5, 32768 + 300, 100, 0), (5, 20000, 60, 3)])
def test_fast_parity_mode_search_run_replayed_by_the_oracle(n, B, S, schedule):
    """HMZ_MODE_FP32X3 through hmz_search_run (configs[1] size and two larger ragged batches; float32 latents, stream
    groups): the network outputs every backup consumed are captured and the C oracle must reproduce visit counts, root
    values and min/max from them bit for bit, over two moves."""
    from muzero_hanoi_b200 import _lib
    from muzero_hanoi_b200.engine import BatchedMCTS, PackedWeights, VecHanoi

    w = PackedWeights(port.make_weights(n, 8), n, _lib.MODE_FP32X3)
    env = VecHanoi(n, 200, B)
    env.random_reset(seed=4)
    rng = np.random.default_rng(12)
    m = BatchedMCTS(0.8, 0.25, S, B, latent_dtype=_lib.LATENT_F32)
    m.store.set_schedule(schedule)
    cap = m.store.enable_capture(S)
    table = port.ucb_table(S + 1)
    for move in range(2):
        noise = rng.dirichlet(np.full(6, 0.25), B)
        uni = rng.random(B)
        mm = m.store.minmax.cpu().numpy().copy()
        cap.fill_(float("nan"))
        action, pi, q, visits = m.run_mcts(w, words=env.words, temperature=1.0, deterministic=False, noise=noise, uniforms=uni)
        torch.cuda.synchronize()
        c = cap.cpu().numpy()
        assert np.isfinite(c).all()
        prior = m.store.root_prior.cpu().numpy()
        assert np.array_equal(prior, port.mix_dirichlet(m.p0.cpu().numpy(), noise))
        o_visits, o_q, _ = cport.search_injected(prior, True, mm, np.ascontiguousarray(c[:, :, 6]), np.ascontiguousarray(c[:, :, :6]),
                                                 np.ascontiguousarray(c[:, :, 7]), 0.8, table)
        assert np.array_equal(visits.cpu().numpy(), o_visits), f"move {move}: visit counts differ from the oracle replay"
        assert np.array_equal(q.cpu().numpy(), o_q)
        assert np.array_equal(m.store.minmax.cpu().numpy(), mm)
        env.step(action.to(torch.uint8), want_obs=False)


@pytest.mark.parametrize("schedule", [0, 64, 128])
def test_full_size_benchmarked_path_replayed_by_the_oracle(schedule):
    """The path bench.py times, at the size it times it: bf16 throughput mode, hmz_search_run (fused backup + select
    kernels, stream groups and programmatic dependent launches for schedule 0; the persistent kernel for schedule 64;
    resident network CTAs fed by ordinary tree launches through release / acquire flags for schedule 128),
    N = 5, 65,536 searches x 100 simulations, two consecutive moves (the second one starts from persisted min/max
    bounds).  The run records the network outputs every backup consumed (hmz_search_t.capture); the C oracle replays
    the tree arithmetic with exactly those outputs injected and must reproduce visit counts, root values (float64) and
    the persistent min/max bit for bit.  An ordering bug between the network kernel and the tree kernel (a backup that
    ran on stale or half-written outputs, a walk that missed a slot update) would show here and nowhere else."""
    from muzero_hanoi_b200 import _lib
    from muzero_hanoi_b200.engine import BatchedMCTS, PackedWeights, VecHanoi

    n, B, S = 5, 65536, 100
    w = PackedWeights(port.make_weights(n, 21), n, _lib.MODE_BF16)
    env = VecHanoi(n, 200, B)
    env.random_reset(seed=2)
    rng = np.random.default_rng(11)
    m = BatchedMCTS(0.8, 0.25, S, B, latent_dtype=_lib.LATENT_BF16)
    m.store.set_schedule(schedule)
    cap = m.store.enable_capture(S)
    table = port.ucb_table(S + 1)
    for move in range(2):
        noise = rng.dirichlet(np.full(6, 0.25), B)
        uni = rng.random(B)
        mm = m.store.minmax.cpu().numpy().copy()
        cap.fill_(float("nan"))
        action, pi, q, visits = m.run_mcts(w, words=env.words, temperature=1.0, deterministic=False, noise=noise, uniforms=uni)
        torch.cuda.synchronize()
        c = cap.cpu().numpy()
        assert np.isfinite(c).all()  # every (simulation, search) cell was written
        prior = m.store.root_prior.cpu().numpy()
        assert np.array_equal(prior, port.mix_dirichlet(m.p0.cpu().numpy(), noise))
        o_visits, o_q, _ = cport.search_injected(prior, True, mm, np.ascontiguousarray(c[:, :, 6]), np.ascontiguousarray(c[:, :, :6]),
                                                 np.ascontiguousarray(c[:, :, 7]), 0.8, table)
        assert np.array_equal(visits.cpu().numpy(), o_visits), f"move {move}: visit counts differ from the oracle replay"
        assert np.array_equal(q.cpu().numpy(), o_q)
        assert np.array_equal(m.store.minmax.cpu().numpy(), mm)
        assert np.array_equal(pi.cpu().numpy(), o_visits / o_visits.sum(1, keepdims=True))
        env.step(action.to(torch.uint8), want_obs=False)


def test_full_size_config3_grouping_invariance_and_properties():
    """BASELINE.json configs[2] size (N = 5, 65,536 searches x 100 simulations, throughput mode): results must
    not depend on how hmz_search_run cuts the batch into concurrent stream groups (the groups only change
    scheduling: separate streams, programmatic dependent launches), and the size-independent invariants of a
    search must hold: visits sum to S, the play policy sums to 1, the root value is a discounted mixture of
    rewards / values and therefore bounded, the sampled action has a visited child."""
    from muzero_hanoi_b200 import _lib
    from muzero_hanoi_b200.engine import BatchedMCTS, PackedWeights, VecHanoi

    n, B, S = 5, 65536, 100
    w = PackedWeights(port.make_weights(n, 21), n, _lib.MODE_BF16)
    env = VecHanoi(n, 200, B)
    env.random_reset(seed=2)
    rng = np.random.default_rng(4)
    noise = torch.from_numpy(rng.dirichlet(np.full(6, 0.25), B)).cuda()
    uni = torch.from_numpy(rng.random(B)).cuda()
    outs = []
    for schedule in SCHEDULES_FULL:
        m = BatchedMCTS(0.8, 0.25, S, B, latent_dtype=_lib.LATENT_BF16)
        m.store.set_schedule(schedule)
        action, pi, q, visits = m.run_mcts(w, words=env.words, temperature=1.0, deterministic=False, noise=noise,
                                           uniforms=uni)
        torch.cuda.synchronize()
        outs.append((action.cpu().numpy(), pi.cpu().numpy(), q.cpu().numpy(), visits.cpu().numpy(),
                     m.store.minmax.cpu().numpy()))
        del m
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b)  # bit for bit, float64 root values and min/max included
    action, pi, q, visits, mm = outs[0]
    assert (visits.sum(1) == S).all() and (visits >= 0).all()
    assert np.allclose(pi.sum(1), 1.0, atol=1e-12) and np.array_equal(pi, visits / visits.sum(1, keepdims=True))
    assert np.isfinite(q).all() and (np.abs(q) <= 278.61 / (1 - 0.8)).all()  # |r|, |v| <= 278.604 (support +-16)
    assert (visits[np.arange(B), action] > 0).all()
    assert (mm[:, 0] <= mm[:, 1]).all()


def test_long_search_beyond_reciprocal_table_bit_exact():
    """S = 14,000 >> the 4,096-entry reciprocal table of the exact-division shortcut: visit counts beyond the table
    take the generic-division fallback (child_score_exact / __ddiv_rn) in both the split-phase kernels (checked
    against the C oracle) and the fused hot-loop kernel (must equal the split phases bit for bit)."""
    from muzero_hanoi_b200.engine import BatchedMCTS, PackedWeights, VecHanoi

    n, B, S = 3, 4, 14000
    weights = PackedWeights(port.make_weights(n, 8), n)
    env = VecHanoi(n, 200, B)
    env.set_state_indices(np.array([0, 9, 13, 25], dtype=np.int32))
    mcts = BatchedMCTS(0.8, 0.0, S, B)
    mm0 = mcts.store.minmax.cpu().numpy().copy()
    p0, r, p, v, depth = _split_search(mcts, weights, env.words, None)
    _, _, q, visits = mcts.root_policy(0.0, True)
    torch.cuda.synchronize()
    assert visits.max().item() > 4096  # some child's count really leaves the table
    mm = mm0.copy()
    o_visits, o_q, o_depth = cport.search_injected(p0.cpu().numpy().astype(np.float64), False, mm, r.cpu().numpy(), p.cpu().numpy(),
                                                   v.cpu().numpy(), 0.8, port.ucb_table(S + 1), want_depth=True)
    assert np.array_equal(visits.cpu().numpy(), o_visits) and np.array_equal(q.cpu().numpy(), o_q)
    assert np.array_equal(depth.cpu().numpy().astype(np.uint16), o_depth)
    assert np.array_equal(mcts.store.minmax.cpu().numpy(), mm)
    fused = BatchedMCTS(0.8, 0.0, S, B)
    _, _, q2, visits2 = fused.run_mcts(weights, words=env.words, temperature=0.0, deterministic=True)
    torch.cuda.synchronize()
    assert np.array_equal(visits2.cpu().numpy(), o_visits) and np.array_equal(q2.cpu().numpy(), o_q)
    assert np.array_equal(fused.store.minmax.cpu().numpy(), mm)


def test_fused_backup_deep_chains_bit_exact():
    """A policy head with one dominant action makes every search a chain that deepens with each simulation: walk
    depths run from 1 past 4 (lane 1's share of the path elements), past 8 (further batches loaded inside the
    backup) and past the 32 recorded levels (parent-link fallback).  The fused hot-loop kernel must equal the C
    oracle bit for bit on visit counts, root values, min/max and — through the split-phase run that records them —
    leaf depths."""
    from muzero_hanoi_b200.engine import BatchedMCTS, PackedWeights, VecHanoi

    n, B, S = 4, 2048, 60
    sd = {k: np.array(v, copy=True) for k, v in port.make_weights(n, 31).items()}
    bias = sd["policy_net.2.bias"]
    bias[:] += np.array([12.0, 0.0, 6.0, 0.0, 0.0, 0.0], dtype=bias.dtype)  # action 0 dominates, action 2 a distant second
    weights = PackedWeights(sd, n)
    env = VecHanoi(n, 200, B)
    env.random_reset(seed=9)
    mcts = BatchedMCTS(0.8, 0.0, S, B)
    mm = mcts.store.minmax.cpu().numpy().copy()
    p0, r, p, v, depth = _split_search(mcts, weights, env.words, None)
    torch.cuda.synchronize()
    o_visits, o_q, o_depth = cport.search_injected(p0.cpu().numpy().astype(np.float64), False, mm, r.cpu().numpy(), p.cpu().numpy(),
                                                   v.cpu().numpy(), 0.8, port.ucb_table(S + 1), want_depth=True)
    d = depth.cpu().numpy().astype(np.uint16)
    assert np.array_equal(d, o_depth)
    assert d.max() > 32 and ((d > 4) & (d <= 8)).any() and ((d > 8) & (d <= 32)).any() and (d <= 4).any()
    for groups_B in (B, 100):  # a ragged last block as well
        fused = BatchedMCTS(0.8, 0.0, S, groups_B)
        sub = VecHanoi(n, 200, groups_B)
        sub.words.copy_(env.words[:groups_B])
        _, _, q2, visits2 = fused.run_mcts(weights, words=sub.words, temperature=0.0, deterministic=True)
        torch.cuda.synchronize()
        assert np.array_equal(visits2.cpu().numpy(), o_visits[:groups_B]) and np.array_equal(q2.cpu().numpy(), o_q[:groups_B])
        assert np.array_equal(fused.store.minmax.cpu().numpy(), mm[:groups_B])


def test_fused_untame_bounds_and_values_take_the_exact_path():
    """The hot loop skips the per-operand range tests of the exact-division shortcut and trusts (a) the persisted
    (min, max) bounds, tested once per launch, and (b) a sticky per-search flag kept by the backup.  (a) bounds far
    outside the tame range (1e-300) must give the oracle's results bit for bit; (b) non-finite network values (NaN
    weights in the value head) must give exactly what the split-phase kernels (per-operand tests) give."""
    from muzero_hanoi_b200.engine import BatchedMCTS, PackedWeights, VecHanoi

    n, B, S = 3, 512, 40
    env = VecHanoi(n, 200, B)
    env.random_reset(seed=3)
    # (a) untame persisted bounds on every second search
    weights = PackedWeights(port.make_weights(n, 5), n)
    mm0 = np.tile(np.array([[np.inf, -np.inf]]), (B, 1))
    mm0[::2] = [1e-300, 3e-300]
    mm0[1::4] = [-2.5, 1e-310]
    mcts = BatchedMCTS(0.8, 0.0, S, B)
    mcts.store.minmax.copy_(torch.from_numpy(mm0))
    p0, r, p, v, depth = _split_search(mcts, weights, env.words, None)
    _, _, q, visits = mcts.root_policy(0.0, True)
    torch.cuda.synchronize()
    mm = mm0.copy()
    o_visits, o_q, _ = cport.search_injected(p0.cpu().numpy().astype(np.float64), False, mm, r.cpu().numpy(), p.cpu().numpy(),
                                             v.cpu().numpy(), 0.8, port.ucb_table(S + 1), want_depth=True)
    assert np.array_equal(visits.cpu().numpy(), o_visits) and np.array_equal(q.cpu().numpy(), o_q)
    fused = BatchedMCTS(0.8, 0.0, S, B)
    fused.store.minmax.copy_(torch.from_numpy(mm0))
    _, _, q2, visits2 = fused.run_mcts(weights, words=env.words, temperature=0.0, deterministic=True)
    torch.cuda.synchronize()
    assert np.array_equal(visits2.cpu().numpy(), o_visits) and np.array_equal(q2.cpu().numpy(), o_q)
    assert np.array_equal(fused.store.minmax.cpu().numpy(), mm)
    # (b) NaN values: fused == split phases, NaNs included
    sd = {k: np.array(x, copy=True) for k, x in port.make_weights(n, 5).items()}
    sd["value_net.2.bias"][:] = np.nan
    bad = PackedWeights(sd, n)
    split = BatchedMCTS(0.8, 0.0, S, B)
    _split_search(split, bad, env.words, None)
    _, _, q3, visits3 = split.root_policy(0.0, True)
    fused = BatchedMCTS(0.8, 0.0, S, B)
    _, _, q4, visits4 = fused.run_mcts(bad, words=env.words, temperature=0.0, deterministic=True)
    torch.cuda.synchronize()
    assert np.array_equal(visits3.cpu().numpy(), visits4.cpu().numpy())
    assert np.array_equal(q3.cpu().numpy(), q4.cpu().numpy(), equal_nan=True)
    assert np.array_equal(split.store.minmax.cpu().numpy(), fused.store.minmax.cpu().numpy(), equal_nan=True)
    assert np.isnan(q4.cpu().numpy()).any()


@pytest.mark.parametrize("n,B,S,bias", [(5, 1, 8, 0.0), (5, 100, 1, 0.0), (3, 777, 30, 0.0), (4, 2048 + 17, 60, 12.0), (5, 8192, 100, 0.0)])
def test_persistent_schedule_equals_serial_launches(n, B, S, bias):
    """hmz_search_t.schedule = HMZ_SCHEDULE_PERSISTENT (one role-specialised kernel per search: tcgen05 MLP CTAs and tree
    CTAs handing 256-search tile pairs to each other through release / acquire counters) against schedule 1 (one launch
    pair per simulation, strictly serial): same arithmetic, so the whole search state must be identical bit for bit —
    node records, latent rows, root values, persistent min/max — over ragged batches (1, 100, 777, 2,065 searches: partial
    tiles, partial warps), 1 to 100 simulations, two consecutive moves, and deep chains (a dominant policy action drives
    walks past the 4 / 8 / 32-level branches of the backup)."""
    from muzero_hanoi_b200 import _lib
    from muzero_hanoi_b200.engine import BatchedMCTS, PackedWeights, VecHanoi

    sd = {k: np.array(v, copy=True) for k, v in port.make_weights(n, 31).items()}
    sd["policy_net.2.bias"][:] += np.array([bias, 0.0, bias / 2, 0.0, 0.0, 0.0], dtype=np.float32)
    w = PackedWeights(sd, n, _lib.MODE_BF16)
    rng = np.random.default_rng(B + S)
    runs = []
    for schedule in (1, _lib.SCHEDULE_PERSISTENT, _lib.SCHEDULE_SERVER, _lib.SCHEDULE_SERVER | 1):
        env = VecHanoi(n, 200, B)
        env.random_reset(seed=9)
        m = BatchedMCTS(0.8, 0.25, S, B, latent_dtype=_lib.LATENT_BF16)
        m.store.set_schedule(schedule)
        rng = np.random.default_rng(B + S)
        out = []
        for move in range(2):
            noise, uni = rng.dirichlet(np.full(6, 0.25), B), rng.random(B)
            action, pi, q, visits = m.run_mcts(w, words=env.words, temperature=1.0, deterministic=False, noise=noise, uniforms=uni)
            torch.cuda.synchronize()
            out.append([t.clone() for t in (action, pi, q, visits, m.store.minmax, m.store.root_W, m.store.nodes,
                                            m.store.latents[:, : S + 1].contiguous())])
            env.step(action.to(torch.uint8), want_obs=False)
        runs.append(out)
    names = ("action", "pi", "root_q", "visits", "minmax", "root_W", "node records", "latents")
    for move in range(2):
        for other, label in ((1, "persistent"), (2, "server (4 groups)"), (3, "server (1 group)")):
            for name, a, b in zip(names, runs[0][move], runs[other][move]):
                assert torch.equal(a, b), f"move {move}: {name} differ between the serial and the {label} schedule"
        assert (runs[1][move][3].sum(1) == S).all()
    if bias:
        depth_proxy = runs[1][1][3].max(1).values  # a dominant action concentrates the visits: chains well past 32 levels
        assert int(depth_proxy.max()) >= S - 6
