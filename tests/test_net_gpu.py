"""GPU parity: MuZeroNet inference kernels (through the C ABI) vs outputs recorded from the
unmodified reference network (tests/golden/net_io.npz).

Tolerances (HMZ_MODE_FP32, north_star: within 1e-5 relative in fp32), all RELATIVE with a small absolute floor:
  latent h and policy p:                       |d| <= 1e-5 * |ref| + 1e-6 (h) / 1e-7 (p)
  reward r and value v (support transform):    |d| <= 1e-5 * |ref| + 2.5e-4
The floor for h is the float32 rounding of normalize_h_state itself ((h - min) / (max - min + 1e-8), networks.py:191-196:
every element is a difference of O(1) numbers, so elements near zero carry an absolute error of a few ulp(1) = 1.2e-7
whatever the kernel does — the smallest element is exactly 0 in both; measured on the B200: <= 4.9e-7 for h, <= 4.5e-8
for p, r and v identical to the reference except for single grid steps, profiles/r02_net_errors.json).  The absolute term for r/v is the float32
granularity of the reference's OWN signed-parabolic evaluation (networks.py:186-189 computes
sqrt(..)/2/eps - 1/2/eps ~ 500.x - 500 in float32, i.e. its outputs live on a ~1.2e-4 grid near zero), so a 1-ulp
difference in the softmax expectation moves the reference's result by one grid step; two grid steps are allowed.
The achieved maxima are recorded by every test (conftest.record_metric -> gpurun_out/test_metrics.jsonl)."""
import numpy as np
import pytest
import torch

from conftest import record_metric
from oracle import port

pytestmark = pytest.mark.gpu
REL, H_ABS, P_ABS, RV_REL, RV_ABS = 1e-5, 1e-6, 1e-7, 1e-5, 2.5e-4
H_TOL = 1e-5  # drop-in surface checks (single rows through the Python shim)


def _close(got, ref, atol):
    return bool(np.all(np.abs(got - ref) <= REL * np.abs(ref) + atol))


def _close_rv(got, ref):
    return np.all(np.abs(got - ref) <= RV_REL * np.abs(ref) + RV_ABS)


def _errs(got, ref):
    """(max absolute error, max relative error over elements with |ref| >= 1e-3)."""
    d = np.abs(np.asarray(got, np.float64) - np.asarray(ref, np.float64))
    big = np.abs(ref) >= 1e-3
    return float(d.max()), float((d[big] / np.abs(ref)[big]).max()) if big.any() else 0.0


def _weights(n, seed, mode=0):
    from muzero_hanoi_b200.engine import PackedWeights

    return PackedWeights(port.make_weights(n, seed), n, mode)


@pytest.mark.parametrize("mode", [0, 2])  # 0: FFMA kernel, 2: HMZ_MODE_FP32X3 (tcgen05, three bf16 parts per operand)
@pytest.mark.parametrize("n", [3, 5, 10])
@pytest.mark.parametrize("count", [96, 33, 1])
def test_recurrent_matches_reference(golden, n, count, mode):
    g = golden("net_io.npz")
    w = _weights(n, int(g[f"n{n}_weight_seed"]), mode)
    h_in = torch.from_numpy(g[f"n{n}_h_in"][:count]).cuda()
    acts = torch.from_numpy(g[f"n{n}_action"][:count].astype(np.uint8)).cuda()
    h = torch.empty(count, 64, device="cuda")
    r, v, p = torch.empty(count, device="cuda"), torch.empty(count, device="cuda"), torch.empty(count, 6, device="cuda")
    w.recurrent(count, latents_in=h_in, in_rows_per_item=1, in_row=None, actions=acts, latents_out=h,
                out_rows_per_item=1, out_row=0, latent_dtype=0, r=r, p=p, v=v)
    torch.cuda.synchronize()
    hh, pp, rr, vv = h.cpu().numpy(), p.cpu().numpy(), r.cpu().numpy(), v.cpu().numpy()
    eh, ep = _errs(hh, g[f"n{n}_h_out"][:count]), _errs(pp, g[f"n{n}_p"][:count])
    er, ev = _errs(rr, g[f"n{n}_r"][:count]), _errs(vv, g[f"n{n}_v"][:count])
    record_metric(f"{'fp32x3' if mode else 'fp32'}_recurrent_n{n}_x{count}", dict(h_abs=eh[0], h_rel=eh[1], p_abs=ep[0], p_rel=ep[1], r_abs=er[0], r_rel=er[1],
                                                        v_abs=ev[0], v_rel=ev[1]))
    assert _close(hh, g[f"n{n}_h_out"][:count], H_ABS)
    assert _close(pp, g[f"n{n}_p"][:count], P_ABS)
    assert _close_rv(rr, g[f"n{n}_r"][:count]) and _close_rv(vv, g[f"n{n}_v"][:count])


def test_fp32x3_many_tiles_against_the_ffma_kernel(golden):
    """HMZ_MODE_FP32X3 at a batch that gives every CTA several 128-row tiles (40,000 + a ragged tail), gathered through
    in_row and scattered to out_row like the search does, float32 and bf16 latent stores: same gate against the FFMA
    kernel's outputs as both have against the reference (the FFMA kernel is pinned to the reference above)."""
    g = golden("net_io.npz")
    n, count, E = 5, 40000 + 77, 3
    sd_seed = int(g[f"n{n}_weight_seed"])
    rng = np.random.default_rng(5)
    src = rng.integers(0, 96, count)
    rows = rng.integers(0, E, count)
    lat = torch.zeros(count, E, 64, device="cuda")
    lat[torch.arange(count), torch.from_numpy(rows)] = torch.from_numpy(g[f"n{n}_h_in"][src]).cuda()
    acts = torch.from_numpy(rng.integers(0, 6, count).astype(np.uint8)).cuda()
    in_row = torch.from_numpy(rows.astype(np.int16)).cuda()
    outs = []
    for mode in (0, 2):
        w = _weights(n, sd_seed, mode)
        out = torch.zeros(count, E + 1, 64, device="cuda")
        r, v, p = torch.empty(count, device="cuda"), torch.empty(count, device="cuda"), torch.empty(count, 6, device="cuda")
        w.recurrent(count, latents_in=lat, in_rows_per_item=E, in_row=in_row, actions=acts, latents_out=out,
                    out_rows_per_item=E + 1, out_row=E, latent_dtype=0, r=r, p=p, v=v)
        torch.cuda.synchronize()
        assert float(out[:, :E].abs().max()) == 0.0
        outs.append((out[:, E].cpu().numpy(), p.cpu().numpy(), r.cpu().numpy(), v.cpu().numpy()))
    (h0, p0, r0, v0), (h2, p2, r2, v2) = outs
    eh, ep, er, ev = _errs(h2, h0), _errs(p2, p0), _errs(r2, r0), _errs(v2, v0)
    record_metric("fp32x3_vs_ffma_n5_x40077", dict(h_abs=eh[0], h_rel=eh[1], p_abs=ep[0], p_rel=ep[1], r_abs=er[0], r_rel=er[1],
                                                   v_abs=ev[0], v_rel=ev[1]))
    assert _close(h2, h0, H_ABS) and _close(p2, p0, P_ABS) and _close_rv(r2, r0) and _close_rv(v2, v0)
    # bf16 latents in and out (the kernel accepts both dtypes): one part is then non-zero
    w = _weights(n, sd_seed, 2)
    outb = torch.zeros(count, 64, dtype=torch.bfloat16, device="cuda")
    r, v, p = torch.empty(count, device="cuda"), torch.empty(count, device="cuda"), torch.empty(count, 6, device="cuda")
    w.recurrent(count, latents_in=lat.to(torch.bfloat16), in_rows_per_item=E, in_row=in_row, actions=acts, latents_out=outb,
                out_rows_per_item=1, out_row=0, latent_dtype=1, r=r, p=p, v=v)
    torch.cuda.synchronize()
    assert np.abs(outb.float().cpu().numpy() - h0).max() <= 2e-2


@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("n", [3, 5, 10])
def test_initial_matches_reference_words_and_obs_paths(golden, n, mode):
    from muzero_hanoi_b200.engine import VecHanoi

    g = golden("net_io.npz")
    count = 96
    w = _weights(n, int(g[f"n{n}_weight_seed"]), mode)  # (mode 2: the FFMA kernel on the blob's embedded float32 copy)
    env = VecHanoi(n, 200, count)
    env.set_state_indices(g[f"n{n}_state_idx"].astype(np.int32))
    outs = []
    for use_words in (True, False):
        h = torch.empty(count, 64, device="cuda")
        p0, v0 = torch.empty(count, 6, device="cuda"), torch.empty(count, device="cuda")
        w.initial(count, words=env.words if use_words else None, obs=None if use_words else env.onehot(),
                  latents_out=h, out_rows_per_item=1, latent_dtype=0, p0=p0, v0=v0)
        torch.cuda.synchronize()
        outs.append((h.cpu().numpy(), p0.cpu().numpy(), v0.cpu().numpy()))
        eh, ep, ev = _errs(outs[-1][0], g[f"n{n}_h0"]), _errs(outs[-1][1], g[f"n{n}_p0"]), _errs(outs[-1][2], g[f"n{n}_v0"])
        record_metric(f"{'fp32x3' if mode else 'fp32'}_initial_n{n}_words{int(use_words)}", dict(h_abs=eh[0], h_rel=eh[1], p_abs=ep[0], p_rel=ep[1], v_abs=ev[0], v_rel=ev[1]))
        assert _close(outs[-1][0], g[f"n{n}_h0"], H_ABS)
        assert _close(outs[-1][1], g[f"n{n}_p0"], P_ABS)
        assert _close_rv(outs[-1][2], g[f"n{n}_v0"])
    if mode == 0:
        for a, b in zip(*outs):  # packed-word and float-observation paths add the same terms in the same order
            assert np.array_equal(a, b)
    # (mode 2: packed words run the tensor-core kernel, float observations the FFMA kernel on the embedded float32 copy;
    #  both are gated against the reference above)


def test_fp32x3_initial_many_tiles_and_record_stride(golden):
    """HMZ_MODE_FP32X3 root inference from packed env words (tcgen05 kernel, one-hot tile built on device) over several
    tiles per CTA, written to record 0 of E-record items: against the FFMA kernel's float32 outputs, same gate."""
    from muzero_hanoi_b200.engine import VecHanoi

    g = golden("net_io.npz")
    n, count, E = 5, 148 * 128 * 2 + 1000 + 3, 3
    seed = int(g[f"n{n}_weight_seed"])
    env = VecHanoi(n, 200, count)
    env.random_reset(seed=8)
    outs = []
    for mode in (0, 2):
        w = _weights(n, seed, mode)
        h = torch.zeros(count, E, 64, device="cuda")
        p0, v0 = torch.empty(count, 6, device="cuda"), torch.empty(count, device="cuda")
        w.initial(count, words=env.words, latents_out=h, out_rows_per_item=E, latent_dtype=0, p0=p0, v0=v0)
        torch.cuda.synchronize()
        assert not bool(h[:, 1:].any())  # only record 0 of each item is written
        outs.append((h[:, 0].cpu().numpy(), p0.cpu().numpy(), v0.cpu().numpy()))
    (h0, p0, v0), (h2, p2, v2) = outs
    eh, ep, ev = _errs(h2, h0), _errs(p2, p0), _errs(v2, v0)
    record_metric("fp32x3_initial_vs_ffma_n5", dict(h_abs=eh[0], h_rel=eh[1], p_abs=ep[0], p_rel=ep[1], v_abs=ev[0], v_rel=ev[1]))
    assert _close(h2, h0, H_ABS) and _close(p2, p0, P_ABS) and _close_rv(v2, v0)


def test_gather_scatter_rows_and_bf16_latent_store(golden):
    """The search-facing addressing: input row = item*E_in + in_row[item], output row = item*E_out + out_row."""
    g = golden("net_io.npz")
    n, count, E = 3, 40, 5
    w = _weights(n, int(g["n3_weight_seed"]))
    rows = np.arange(count) % E
    lat = torch.zeros(count, E, 64, device="cuda")
    lat[torch.arange(count), torch.from_numpy(rows)] = torch.from_numpy(g["n3_h_in"][:count]).cuda()
    out = torch.zeros(count, E + 1, 64, device="cuda")
    acts = torch.from_numpy(g["n3_action"][:count].astype(np.uint8)).cuda()
    r, v, p = torch.empty(count, device="cuda"), torch.empty(count, device="cuda"), torch.empty(count, 6, device="cuda")
    w.recurrent(count, latents_in=lat, in_rows_per_item=E, in_row=torch.from_numpy(rows.astype(np.int16)).cuda(),
                actions=acts, latents_out=out, out_rows_per_item=E + 1, out_row=E, latent_dtype=0, r=r, p=p, v=v)
    torch.cuda.synchronize()
    assert np.abs(out[:, E].cpu().numpy() - g["n3_h_out"][:count]).max() <= H_TOL
    assert float(out[:, :E].abs().max()) == 0.0
    outb = torch.zeros(count, 64, dtype=torch.bfloat16, device="cuda")
    w.recurrent(count, latents_in=lat.to(torch.bfloat16), in_rows_per_item=E,
                in_row=torch.from_numpy(rows.astype(np.int16)).cuda(), actions=acts, latents_out=outb,
                out_rows_per_item=1, out_row=0, latent_dtype=1, r=r, p=p, v=v)
    torch.cuda.synchronize()
    assert np.abs(outb.float().cpu().numpy() - g["n3_h_out"][:count]).max() <= 2e-2


def test_dropin_muzeronet_surface(golden):
    from muzero_hanoi_b200.networks import MuZeroNet

    g = golden("net_io.npz")
    net = MuZeroNet(rpr_input_s=9, action_s=6, lr=0.002, device="cpu", TD_return=True)
    want_keys = [f"{m}.{i}.{k}" for m in ("representation_net", "dynamic_net", "rwd_net", "policy_net", "value_net")
                 for i in (0, 2) for k in ("weight", "bias")]
    assert list(net.state_dict().keys()) == want_keys
    net.load_state_dict({k: torch.from_numpy(v) for k, v in port.make_weights(3, 0).items()})
    a1 = torch.zeros(6)
    a1[int(g["n3_action"][0])] = 1.0
    h, r, p, v = net.recurrent_inference(torch.from_numpy(g["n3_h_in"][0]), a1)
    assert h.dtype == np.float32 and h.shape == (64,) and p.dtype == np.float32 and p.shape == (6,)
    assert isinstance(r, float) and isinstance(v, float)
    assert np.abs(h - g["n3_h_out"][0]).max() <= H_TOL and _close_rv(np.float32(r), g["n3_r"][0])
    obs = port.one_hot(port.index_to_state(int(g["n3_state_idx"][0]), 3))
    h0, r0, p0, v0 = net.initial_inference(torch.from_numpy(obs).to(torch.float32))
    assert r0 == 0.0 and np.abs(p0 - g["n3_p0"][0]).max() <= H_TOL
    h0b, _, _, _ = net.initial_inference(torch.from_numpy(obs).to(torch.float32).reshape(1, 9))  # batched [1, 9] form
    assert np.array_equal(h0, h0b)
    # lesion: reset_param changes the weights, the packed blob follows (acting_ablations.py:29-45)
    torch.manual_seed(0)
    net.policy_net.apply(net.reset_param)
    _, _, p1, _ = net.initial_inference(torch.from_numpy(obs).to(torch.float32))
    assert not np.array_equal(p0, p1)
    with torch.no_grad():  # differentiable torch forms stay consistent with the kernels
        hh = net.represent(torch.from_numpy(obs).to(torch.float32))
    assert np.abs(hh.numpy() - h0).max() <= H_TOL


# ------------------------------------------------------------- tensor-core (bf16) path
BF16_TOL = 2e-2  # north_star: network outputs within 2e-2 in bf16 (relative to each output's scale)


def _bf16_close(got, ref, scale):
    return np.abs(got - ref).max() <= BF16_TOL * scale


@pytest.mark.parametrize("n", [3, 5])
@pytest.mark.parametrize("count,latent_dtype", [(96, 0), (1, 0), (300, 1)])
def test_recurrent_tcgen05_bf16_matches_reference(golden, n, count, latent_dtype):
    """HMZ_MODE_BF16 (tcgen05 + TMEM + TMA bulk): h, p within 2e-2 absolute (unit scale), r and v
    within 2e-2 of max(1, |ref|) — bf16 operands, fp32 accumulation."""
    g = golden("net_io.npz")
    w = _weights(n, int(g[f"n{n}_weight_seed"]), mode=1)
    idx = np.arange(count) % 96
    h_in = torch.from_numpy(g[f"n{n}_h_in"][idx]).cuda()
    if latent_dtype == 1:
        h_in = h_in.to(torch.bfloat16)
    acts = torch.from_numpy(g[f"n{n}_action"][idx].astype(np.uint8)).cuda()
    h = torch.empty(count, 64, device="cuda", dtype=torch.bfloat16 if latent_dtype else torch.float32)
    r, v, p = torch.empty(count, device="cuda"), torch.empty(count, device="cuda"), torch.empty(count, 6, device="cuda")
    w.recurrent(count, latents_in=h_in, in_rows_per_item=1, in_row=None, actions=acts, latents_out=h,
                out_rows_per_item=1, out_row=0, latent_dtype=latent_dtype, r=r, p=p, v=v)
    torch.cuda.synchronize()
    hh, rr, vv, pp = h.float().cpu().numpy(), r.cpu().numpy(), v.cpu().numpy(), p.cpu().numpy()
    print("max |dh| %.4f |dp| %.4f |dr| %.4f |dv| %.4f" % (
        np.abs(hh - g[f"n{n}_h_out"][idx]).max(), np.abs(pp - g[f"n{n}_p"][idx]).max(),
        np.abs(rr - g[f"n{n}_r"][idx]).max(), np.abs(vv - g[f"n{n}_v"][idx]).max()))
    assert _bf16_close(hh, g[f"n{n}_h_out"][idx], 1.0)
    assert _bf16_close(pp, g[f"n{n}_p"][idx], 1.0)
    assert np.all(np.abs(rr - g[f"n{n}_r"][idx]) <= BF16_TOL * np.maximum(1.0, np.abs(g[f"n{n}_r"][idx])))
    assert np.all(np.abs(vv - g[f"n{n}_v"][idx]) <= BF16_TOL * np.maximum(1.0, np.abs(g[f"n{n}_v"][idx])))


@pytest.mark.parametrize("n", [3, 5, 10])
@pytest.mark.parametrize("latent_dtype", [0, 1])
def test_initial_tcgen05_bf16_matches_reference(golden, n, latent_dtype):
    """HMZ_MODE_BF16 root inference from packed env words (tcgen05 kernel, one-hot tile built on device):
    h0, p0 within 2e-2 absolute, v0 within 2e-2 of max(1, |ref|)."""
    from muzero_hanoi_b200.engine import VecHanoi

    g = golden("net_io.npz")
    count, E = 96, 3
    w = _weights(n, int(g[f"n{n}_weight_seed"]), mode=1)
    env = VecHanoi(n, 200, count)
    env.set_state_indices(g[f"n{n}_state_idx"].astype(np.int32))
    h = torch.zeros(count, E, 64, device="cuda", dtype=torch.bfloat16 if latent_dtype else torch.float32)
    p0, v0 = torch.empty(count, 6, device="cuda"), torch.empty(count, device="cuda")
    w.initial(count, words=env.words, latents_out=h, out_rows_per_item=E, latent_dtype=latent_dtype, p0=p0, v0=v0)
    torch.cuda.synchronize()
    hh = h.float().cpu().numpy()
    assert _bf16_close(hh[:, 0], g[f"n{n}_h0"], 1.0) and not hh[:, 1:].any()  # only record 0 of each item is written
    assert _bf16_close(p0.cpu().numpy(), g[f"n{n}_p0"], 1.0)
    vv = v0.cpu().numpy()
    assert np.all(np.abs(vv - g[f"n{n}_v0"]) <= BF16_TOL * np.maximum(1.0, np.abs(g[f"n{n}_v0"])))


def test_bf16_mode_search_runs_and_agrees_with_fp32_mode():
    """Throughput mode end to end: same searches in bf16 and fp32 modes give visit counts that sum to
    S and root policies that are close (bf16 rounding moves a few visits, not the search)."""
    from muzero_hanoi_b200.engine import BatchedMCTS, PackedWeights, VecHanoi

    n, B, S = 5, 1000, 50
    sd = port.make_weights(n, 4)
    env = VecHanoi(n, 200, B)
    env.random_reset(seed=9)
    out = []
    for mode, ldt in ((0, 0), (1, 1)):
        m = BatchedMCTS(0.8, 0.0, S, B, latent_dtype=ldt)
        _, pi, q, visits = m.run_mcts(PackedWeights(sd, n, mode), words=env.words, temperature=1.0, deterministic=True)
        torch.cuda.synchronize()
        assert (visits.sum(1) == S).all()
        out.append((pi.cpu().numpy(), q.cpu().numpy()))
    moved = np.abs(out[0][0] - out[1][0]).sum(1).mean() / 2
    record_metric("bf16_vs_fp32_search_n5_1000x50", dict(mean_fraction_of_visits_moved=moved,
                                                         max_abs_root_q_diff=np.abs(out[0][1] - out[1][1]).max()))
    assert moved < 0.10
