"""GPU parity: tree-store kernels (select / expand+backup / root policy) vs the golden search
traces recorded from the unmodified reference.  Network outputs are injected (north_star parity
mode), noise and sampling uniforms are inputs, ties break to the lowest index.  Bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
FIXTURES = ["n3_s50_noise_t1", "n3_s25_nonoise_t0", "n3_s25_det", "n4_s200_lesion_t0", "n5_s100_noise_t05",
            "n3_s50_noise_t01",   # T = 0.1 -> exponent 5 (utils.py:89-96 for episodes >= 750; np.power(int64, 5.0), mcts.py:170-174)
            "n4_s60_noise_t04"]   # T = 0.4 -> exponent 2.5: NumPy's vectorised pow (host table through the ABI)
INF = float("inf")


def _check_against_fixture(g, mcts, ks, paths, depths):
    S = int(g["S"])
    act, pi, q, visits = mcts.root_policy(float(g["temperature"]), bool(g["deterministic"]),
                                          uniforms=g["uniform"][ks])
    torch.cuda.synchronize()
    assert np.array_equal(visits.cpu().numpy(), g["child_N"][ks])
    assert np.array_equal(q.cpu().numpy(), g["root_q"][ks])  # float64, bit for bit
    ex = max(1.0, min(5.0, 1.0 / float(g["temperature"]))) if float(g["temperature"]) > 0 else 1.0
    if ex == int(ex):
        assert np.array_equal(pi.cpu().numpy(), g["pi"][ks])
    else:
        # NumPy's vectorised pow is not correctly rounded, differs between CPUs and — its SIMD body and scalar head / tail
        # round differently — even with the memory alignment of the reference's 6-element array: a non-integer power is
        # pinned to the last bit or two only (measured on the B200 box: NumPy disagrees with itself at that level)
        assert np.allclose(pi.cpu().numpy(), g["pi"][ks], rtol=1e-15, atol=0.0)
    assert np.array_equal(act.cpu().numpy(), g["action"][ks])
    mm = mcts.store.minmax.cpu().numpy()
    assert np.array_equal(mm[:, 0], g["mm_min"][ks]) and np.array_equal(mm[:, 1], g["mm_max"][ks])
    if paths is not None:
        paths, depths = paths.cpu().numpy(), depths.cpu().numpy()
        for j, k in enumerate(ks):
            assert np.array_equal(depths[:, j], g["depth"][k])
            D = g["path"].shape[2]
            assert np.array_equal(paths[:, j, :D], g["path"][k][:, :D])
            assert (paths[:, j, D:] == 255).all()


@pytest.mark.parametrize("name", FIXTURES)
def test_batched_injected_search_matches_reference(golden, name):
    """All K recorded searches of a fixture run side by side as one batch, each slot's MinMaxStats
    preloaded with the state the reference had before that search."""
    from muzero_hanoi_b200.engine import BatchedMCTS

    g = golden(f"search_{name}.npz")
    S, K = int(g["S"]), int(g["K"])
    ks = np.arange(K)
    mcts = BatchedMCTS(float(g["discount"]), float(g["alpha"]), S, K)
    mm0 = np.stack([np.concatenate([[INF], g["mm_min"][:-1]]), np.concatenate([[-INF], g["mm_max"][:-1]])], 1)
    mcts.store.minmax.copy_(torch.from_numpy(mm0))
    r = np.ascontiguousarray(g["r"].T)
    v = np.ascontiguousarray(g["v"].T)
    p = np.ascontiguousarray(g["p"].transpose(1, 0, 2))
    paths, depths = mcts.run_injected(g["prior"], bool(g["prior_is_f64"]), r, p, v, want_paths=True)
    _check_against_fixture(g, mcts, ks, paths, depths)


@pytest.mark.parametrize("name", ["n3_s50_noise_t1", "n4_s200_lesion_t0"])
def test_minmax_persists_across_moves(golden, name):
    """One search slot replays the K consecutive moves of the reference episode: the (min, max) of
    MinMaxStats must carry over from move to move exactly as on the reference's MCTS object."""
    from muzero_hanoi_b200.engine import BatchedMCTS

    g = golden(f"search_{name}.npz")
    S, K = int(g["S"]), int(g["K"])
    mcts = BatchedMCTS(float(g["discount"]), float(g["alpha"]), S, 1)
    for k in range(K):
        mcts.run_injected(g["prior"][k:k + 1], bool(g["prior_is_f64"]), g["r"][k][:, None], g["p"][k][:, None, :],
                          g["v"][k][:, None])
        _check_against_fixture(g, mcts, np.array([k]), None, None)


@pytest.mark.parametrize("name", ["n3_s50_noise_t1", "n3_s25_det", "n4_s200_lesion_t0"])
def test_return_latent_actions_equals_reference_last_path(golden, name):
    """MCTS.return_latent_actions (MCTS/mcts.py:128-130; filled at :79,84-86) = the actions selected along the LAST
    simulation's path.  Injected mode, so the tree is the reference's tree: the drop-in's list must equal the golden
    path of simulation S - 1 exactly (values, not just shape)."""
    from muzero_hanoi_b200.engine import BatchedMCTS
    from muzero_hanoi_b200.MCTS.mcts import MCTS

    g = golden(f"search_{name}.npz")
    S, K = int(g["S"]), int(g["K"])
    eng = BatchedMCTS(float(g["discount"]), float(g["alpha"]), S, 1)
    shim = MCTS(float(g["discount"]), float(g["alpha"]), S, 1, "cpu")
    shim._engine = eng
    for k in range(K):
        eng.run_injected(g["prior"][k:k + 1], bool(g["prior_is_f64"]), g["r"][k][:, None], g["p"][k][:, None, :], g["v"][k][:, None])
        want = g["path"][k][S - 1, : int(g["depth"][k][S - 1])]
        got = shim.return_latent_actions()
        assert [int(t.item()) for t in got] == want.tolist()
        assert all(t.dtype == torch.long and t.shape == (1,) for t in got) and got is shim.latent_actions


def test_play_policy_powers_device_vs_numpy():
    """generate_play_policy (MCTS/mcts.py:154-176) over every temperature of the reference's schedule and a few others:
    integer exponents (1 .. 5) must reproduce NumPy bit for bit, non-integer ones to 1e-15 relative (NumPy's own pow is
    not reproducible beyond that, see engine.BatchedMCTS._pow_table); a caller-supplied pow_table is used verbatim."""
    from muzero_hanoi_b200.engine import BatchedMCTS
    from oracle import port

    S, B = 1500, 64  # 1500 ** 5 < 2 ** 53: the largest counts for which exponent 5 is exact
    rng = np.random.default_rng(5)
    mcts = BatchedMCTS(0.8, 0.0, S, B)
    rec = mcts.store.nodes.view(torch.int16).reshape(B, S + 1, 64)
    counts = rng.multinomial(S, rng.dirichlet(np.full(6, 0.3), B)).astype(np.int32)  # [B, 6], rows sum to S
    counts[0] = [S, 0, 0, 0, 0, 0]
    for a in range(6):  # child a lives in half a // 3, slot a % 3; N is the uint16 at byte 12 of the 16-byte slot
        rec[:, 0, (a // 3) * 32 + (a % 3) * 8 + 6] = torch.from_numpy(counts[:, a].astype(np.int16)).cuda()
    for T in (1.0, 0.5, 0.25, 0.2, 0.1, 0.05, 0.4, 0.3, 0.7, 0.0):
        u = rng.random(B)
        act, pi, _, visits = mcts.root_policy(T, False, uniforms=u)
        torch.cuda.synchronize()
        assert np.array_equal(visits.cpu().numpy(), counts)
        want = np.stack([port.play_policy(c, T) for c in counts])
        ex = max(1.0, min(5.0, 1.0 / T)) if T > 0 else 1.0
        if ex == int(ex):
            assert np.array_equal(pi.cpu().numpy(), want), f"T={T}"
            assert act.cpu().numpy().tolist() == [port.sample_action(want[i], u[i]) for i in range(B)]
        else:
            assert np.allclose(pi.cpu().numpy(), want, rtol=1e-15, atol=0.0), f"T={T}"
    # the pow_table hook: powers supplied by the caller are used as they are (here: n -> n + 1 at every position)
    mcts.pow_table = torch.from_numpy(np.repeat(np.arange(1, S + 2, dtype=np.float64)[:, None], 6, axis=1).copy()).cuda()
    _, pi, _, _ = mcts.root_policy(0.4, False, uniforms=rng.random(B))
    torch.cuda.synchronize()
    assert np.array_equal(pi.cpu().numpy(), (counts + 1.0) / (counts + 1.0).sum(1, keepdims=True))


def test_ragged_batch_and_record_layout(golden):
    """Batch sizes that do not fill a warp segment / block, and the node-record layout itself."""
    from muzero_hanoi_b200.engine import BatchedMCTS

    g = golden("search_n3_s25_nonoise_t0.npz")
    S = int(g["S"])
    for B in (1, 3, 5, 33, 67):
        ks = np.arange(B) % int(g["K"])
        mcts = BatchedMCTS(float(g["discount"]), 0.0, S, B)
        mm0 = np.stack([np.concatenate([[INF], g["mm_min"][:-1]]), np.concatenate([[-INF], g["mm_max"][:-1]])], 1)[ks]
        mcts.store.minmax.copy_(torch.from_numpy(mm0))
        mcts.run_injected(g["prior"][ks], False, np.ascontiguousarray(g["r"][ks].T),
                          np.ascontiguousarray(g["p"][ks].transpose(1, 0, 2)), np.ascontiguousarray(g["v"][ks].T))
        _check_against_fixture(g, mcts, ks, None, None)
    rec = mcts.store.records()
    assert (rec["N"][:, 0].sum(1) == S).all()  # sum of root child visits == n_simulations
    assert (rec["parent"][:, 0] == 0).all()
    for b in range(3):
        for e in range(1, S + 1):
            pe, pa = rec["parent"][b, e], rec["parent_action"][b, e]
            assert rec["child"][b, pe, pa] == e  # parent/child links are mutual
            assert np.array_equal(rec["prior"][b, e], g["p"][ks[b]][e - 1])
            assert rec["rwd"][b, pe, pa] == g["r"][ks[b]][e - 1]
            assert rec["N"][b, e].sum() <= rec["N"][b, pe, pa]  # a node's visits bound its children's


def test_empty_batch_and_bad_arguments():
    from muzero_hanoi_b200 import _lib
    from muzero_hanoi_b200.engine import BatchedMCTS

    mcts = BatchedMCTS(0.8, 0.0, 4, 2)
    with pytest.raises(ValueError, match="temperature"):
        mcts.root_policy(1.5, True)
    with pytest.raises(_lib.HmzError):
        mcts.select(10)  # simulation index beyond the record capacity
    empty = BatchedMCTS(0.8, 0.0, 4, 0)
    empty.store.reset_minmax()
