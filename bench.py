#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native MuZero/Hanoi acting engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the config the metric's targets are quoted on): Tower of Hanoi
with 5 disks, 65,536 parallel self-play games PER GPU x 100 MCTS simulations per move, random-init
h/g/f networks, Dirichlet root noise, temperature 1.  A "step" is one move of every game:
root inference -> 100 x (select -> g+f MLP -> expand+backup) -> root policy -> env step.
`value` = MCTS simulations / s over all GPUs (weak scaling: games per GPU fixed), state resident
in HBM; `e2e` = the same step with the env words coming from pinned host memory and the move's
records (state, action, reward, flags, visits, root value) read back to the host every step.
One JSON line on stdout (rank 0); progress goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_DISKS, GAMES_PER_GPU, N_SIMS, MAX_STEPS = 5, 65536, 100, 200
DISCOUNT, ALPHA, EPS, TEMPERATURE = 0.8, 0.25, 0.25, 1.0
FLOP_PER_SIM = 203_776  # 101,888 MAC of g + reward/policy/value heads, un-padded
ENV_BYTES_PER_STEP = 14  # word r/w 8 + action 1 + reward 4 + flags 1
METRIC, UNIT = "mcts_simulations_per_second", "sims/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, for one) print to
    file descriptor 1 from C, so fd 1 is pointed at stderr for the whole run and the JSON line is written
    to a private duplicate of the original stdout at the end."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_tflops_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    """One host process of the reference-style CPU path: the oracle port (numpy tree + one-row
    float32 torch network, oracle/port.py) playing Hanoi self-play moves for `budget_s` seconds
    or `n_moves` moves."""
    seed, n_moves, budget_s = args
    import numpy as np
    import torch

    from oracle import port

    torch.set_num_threads(1)
    rng = np.random.default_rng(seed)
    net = port.PortNet(port.make_weights(N_DISKS, 0))
    env = port.PortHanoi(N_DISKS, MAX_STEPS)
    mm = port.MinMax()
    obs = env.reset()
    sims = moves = 0
    t0 = time.perf_counter()
    while (n_moves is None or moves < n_moves) and (budget_s is None or time.perf_counter() - t0 < budget_s):
        a, _, _, _, _ = port.run_mcts_port(obs, net, port.PortSearch(DISCOUNT, N_SIMS, mm), TEMPERATURE, False,
                                           alpha=ALPHA, noise=rng.dirichlet(np.full(6, ALPHA)), u=rng.random())
        obs, _, done, _ = env.step(a)
        if done:
            obs = env.reset()
        sims += N_SIMS
        moves += 1
    return sims, moves, time.perf_counter() - t0


def cpu_reference_rate(n_procs, n_moves=None, budget_s=None, pool=None):
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(n_procs)
    try:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(1000 + i, n_moves, budget_s) for i in range(n_procs)])
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.close()
            pool.join()
    sims = sum(r[0] for r in res)
    moves = sum(r[1] for r in res)
    return sims / wall, moves / wall, wall, sims


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself is pure
    Python and cannot travel to the box) on all host cores, same metric / unit / config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    moves_per_step = 2
    pool = mp.get_context("fork").Pool(cores)
    try:
        for _ in range(args.warmup):
            cpu_reference_rate(cores, n_moves=1, pool=pool)
        t0 = time.perf_counter()
        sims = 0
        for _ in range(args.steps):
            _, _, _, s = cpu_reference_rate(cores, n_moves=moves_per_step, pool=pool)
            sims += s
        wall = time.perf_counter() - t0
    finally:
        pool.close()
        pool.join()
    value = sims / wall
    sample = f"{cores} processes x {moves_per_step} moves x {N_SIMS} sims per step (N={N_DISKS}), oracle port of the reference"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / max(1, args.steps) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "env_steps_per_second": value / N_SIMS,
    }
    emit(line)


# ------------------------------------------------------------------------------------ ours
def measure_extras(dev, weights, peaks, timed):
    """SURVEY §8f rows 1-3 at BASELINE sizes: episode post-processing kernels, replay-ring insertion and the
    batched acting-evaluation harness.  Synthetic finished episodes (random lengths, flags, visit counts)."""
    import numpy as np
    import torch

    from muzero_hanoi_b200 import _lib, acting
    from muzero_hanoi_b200.engine import PackedWeights
    from muzero_hanoi_b200.replay import EpisodeStore, ReplayRing
    from oracle import port

    B, T, n = GAMES_PER_GPU, MAX_STEPS, N_DISKS
    st = EpisodeStore(B, T, n, dev)
    g = torch.Generator(device=dev).manual_seed(0)
    st.ep_len.copy_(torch.randint(1, T + 1, (B,), generator=g, device=dev, dtype=torch.int32))
    st.flags.copy_((torch.rand(T, B, generator=g, device=dev) < 0.25).to(torch.uint8) * 2)
    last = (st.ep_len.long() - 1).clamp(min=0)
    st.flags[last, torch.arange(B, device=dev)] = 5  # every episode ends solved: all of them enter the ring
    st.root_q.copy_(torch.randn(T, B, generator=g, device=dev, dtype=torch.float64) * 20)
    st.visits.copy_(torch.randint(1, 40, (T, B, 6), generator=g, device=dev, dtype=torch.int16))
    st.state.copy_(torch.randint(0, 1 << (2 * n), (T, B), generator=g, device=dev, dtype=torch.int32))
    steps = int(st.ep_len.sum().item())
    ms_ret = timed(lambda: st.post_process(10, DISCOUNT), 10) / 10
    ms_mc = timed(lambda: st.post_process_mc(DISCOUNT), 10) / 10
    ring = ReplayRing(steps + 1024, 5, 3 * n, 6, dev)

    def add():
        ring.ptr, ring.is_full = 0, False
        ring.add_episodes(st, temperature=TEMPERATURE, only_solved=True)
    add()
    ms_add = timed(add, 5) / 5
    row_bytes = 4 * 3 * n + 5 * 4 + 5 * 8 + 5 * 6 * 4 + 5 * 4 + 4
    out = {"episodes": B, "transitions": steps, "t_max": T,
           "n_step_returns": {"ms": ms_ret, "transitions_per_s": steps / (ms_ret * 1e-3),
                              "gbs": steps * 21 / (ms_ret * 1e-3) / 1e9, "algorithmic": "21 B / transition (flags 1, root_q 8, return 8, priority 4)"},
           "mc_returns": {"ms": ms_mc, "transitions_per_s": steps / (ms_mc * 1e-3)},
           "replay_add": {"ms": ms_add, "rows_per_s": steps / (ms_add * 1e-3), "gbs": steps * (row_bytes + 30) / (ms_add * 1e-3) / 1e9,
                          "frac_of_hbm": steps * (row_bytes + 30) / (ms_add * 1e-3) / 1e9 / peaks["hbm_gbs"],
                          "algorithmic": f"{row_bytes} B row written + 30 B episode record read per transition; includes the "
                                         "row-assignment scan and one scalar read-back"}}
    # acting harness at BASELINE.json configs[4] semantics (N=4, random starts, S=200, T=0), reduced episode count
    n4, eps, sims = 4, 4096, 200
    w4 = PackedWeights(port.make_weights(n4, 7), n4, _lib.MODE_BF16, dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, det = acting.get_results(w4, n4, MAX_STEPS, eps, [sims], 0.0, seed=1, return_details=True)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    moves = int(det[0]["steps"].max())
    # learner step (Muzero._update + Adam) at the reference's default batch (TrainingConfig: batch 256, unroll 5)
    from muzero_hanoi_b200.learner import Learner
    from muzero_hanoi_b200.networks import MuZeroNet

    Bl, Kl = 256, 5
    rng = np.random.default_rng(3)
    sd = port.make_weights(n, 9)
    batch = (np.stack([port.one_hot(port.index_to_state(int(i), n)) for i in rng.integers(0, 3 ** n, Bl)]).astype(np.float32),
             rng.choice(np.array([0.0, 100.0, -0.1], np.float32), size=(Bl, Kl)).astype(np.float32),
             rng.integers(0, 6, (Bl, Kl)).astype(np.int64), rng.dirichlet(np.ones(6), size=(Bl, Kl)).astype(np.float32),
             rng.normal(0, 20, (Bl, Kl)).astype(np.float32), rng.uniform(0.2, 1.0, Bl).astype(np.float32))
    ln = Learner(sd, n, Kl, device=dev)
    dbatch = [torch.as_tensor(x, device=dev) for x in batch]
    ln.update(*dbatch)
    ms_upd = timed(lambda: ln.update(*dbatch), 20) / 20
    # the reference's own step: torch CPU autograd + Adam through the same modules (all host threads torch wants)
    import torch.nn.functional as F
    net = MuZeroNet(3 * n, 6, 0.002, "cpu", TD_return=True)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    tb = [torch.from_numpy(x) for x in batch]

    def ref_update():
        h = net.represent(tb[0])
        loss = 0
        for t in range(Kl):
            pl, pv = net.prediction(h)
            h, pr = net.dynamics(h, F.one_hot(tb[2][:, t], 6).to(torch.long))
            h.register_hook(lambda grad: grad * 0.5)
            loss = loss + F.mse_loss(pv.squeeze(), tb[4][:, t], reduction="none") + F.mse_loss(pr.squeeze(), tb[1][:, t], reduction="none") \
                + F.cross_entropy(pl, tb[3][:, t], reduction="none")
        loss = (loss * tb[5]).mean()
        loss.register_hook(lambda grad: grad * (1 / Kl))
        net.update(loss)
    ref_update()
    t0 = time.perf_counter()
    for _ in range(10):
        ref_update()
    ms_ref = (time.perf_counter() - t0) / 10 * 1e3
    out["learner_step"] = {"workload": f"Muzero._update batch {Bl} x unroll {Kl}, N={n} (TrainingConfig defaults)", "ms": ms_upd,
                           "updates_per_s": 1e3 / ms_upd, "cpu_torch_ms": ms_ref, "cpu_torch_threads": torch.get_num_threads(),
                           "note": "includes the host-side loss read-back of every update"}
    out["acting_harness"] = {"workload": f"hanoi{n4}_{eps}episodes_x{sims}sims_T0 (BASELINE.json configs[4] semantics)",
                             "wall_s": wall, "moves_played": moves, "sims_per_s": eps * sims * moves / wall,
                             "mean_error": float(det[0]["errors"].mean()), "episodes_per_s": eps / wall}
    return out


def workload_config(args, world):
    return {
        "workload": f"hanoi{N_DISKS}_selfplay_{GAMES_PER_GPU}games_per_gpu_x{N_SIMS}sims (BASELINE.json configs[2])",
        "n_disks": N_DISKS, "games_per_gpu": args.games, "global_games": args.games * world, "n_simulations": args.sims,
        "max_steps": MAX_STEPS, "discount": DISCOUNT, "dirichlet_alpha": ALPHA, "temperature": TEMPERATURE,
        "mode": args.mode, "search_groups": args.groups, "parallelism": f"games sharded over {world} GPU(s), NCCL all-gather of move records only",
        "cache": "working set (tree + latents ~2.5 GB/GPU) is larger than the 126 MB L2; no L2 flush needed",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for (t, r) in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.03]
        if not inside:  # region shorter than one sampling period: take the samples closest to it
            inside = [r for (_, r) in self.rows[-3:]]
        for r in inside:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_ours(args):
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    # CPU baseline first (rank 0, N=1 only): fork()ing is only safe before CUDA is initialised.
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    cpu_baseline = None
    if world_env == 1 and args.gpus == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        log(f"[bench] CPU baseline: oracle port on {cores} processes for ~{args.cpu_seconds:.0f} s ...")
        rate, moves_rate, wall, sims = cpu_reference_rate(cores, budget_s=args.cpu_seconds)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{cores} processes x {wall:.1f} s of N={N_DISKS}, S={N_SIMS} self-play moves "
                                  f"({sims} simulations) through oracle/port.py (numpy tree + 1-row fp32 torch net)",
                        "env_steps_per_second": moves_rate}
        log(f"[bench] CPU baseline: {rate:.1f} sims/s on {cores} cores")

    from muzero_hanoi_b200 import _lib
    from muzero_hanoi_b200 import dist as hdist
    from muzero_hanoi_b200.engine import PackedWeights, SelfPlay, VecHanoi
    from muzero_hanoi_b200.networks import MuZeroNet

    rank, world, local = hdist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    mode = _lib.MODE_BF16 if args.mode == "bf16" else _lib.MODE_FP32
    latent_dtype = _lib.LATENT_BF16 if args.mode == "bf16" else _lib.LATENT_F32
    torch.manual_seed(0)
    net = MuZeroNet(3 * N_DISKS, 6, 0.002, "cpu", TD_return=True)  # random-init h / g / f
    weights = PackedWeights(net.state_dict(), N_DISKS, mode, dev)
    B, S = args.games, args.sims
    sp = SelfPlay(N_DISKS, MAX_STEPS, B, S, weights, DISCOUNT, ALPHA, EPS, TEMPERATURE, seed=1234 + rank,
                  ring_slots=4, device=dev, latent_dtype=latent_dtype)
    sp.mcts.store.set_schedule(args.groups)
    gather_buf = torch.empty(world * B, 26, dtype=torch.uint8, device=dev) if world > 1 else None

    def step():
        t = sp.move()
        if world > 1:
            hdist.all_gather_records(sp.slot(t), gather_buf)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started before the warm-up so that nvidia-smi is already streaming when timing begins
    log(f"[bench] rank {rank}/{world}: warm-up {args.warmup} steps (B={B}, S={S}, mode={args.mode})")
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = lib.hmz_launch_count()
    sampler.mark_begin()
    ms = timed(step, args.steps)
    sampler.mark_end()
    launches = lib.hmz_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    sims_per_s = world * B * S * args.steps / (ms * 1e-3)
    log(f"[bench] {ms / args.steps:.2f} ms/step -> {sims_per_s:.3e} sims/s")

    # ---- per-kernel device time over the same steps (CUDA-event pairs on the launch stream)
    barrier()
    sp.mcts.store.set_schedule(1)  # serial launches: clean, non-overlapped per-kernel durations
    _lib.check(lib.hmz_prof_begin())
    for _ in range(args.steps):
        step()
    ms_cls = (C.c_double * 8)()
    n_cls = (C.c_int64 * 8)()
    _lib.check(lib.hmz_prof_end(ms_cls, n_cls))
    sp.mcts.store.set_schedule(args.groups)
    # "backup_select" = the fused expansion + backup(sim) + selection(sim+1) kernel; "select" = the first selection of a move
    names = ["env_step", "select", "net_recurrent", "backup_select", "net_initial", "root_policy", "other", "-"]
    kern = {names[i]: {"ms_total": ms_cls[i], "launches": int(n_cls[i]),
                       "us_per_launch": (ms_cls[i] / n_cls[i] * 1e3) if n_cls[i] else None} for i in range(7)}
    total_kernel_ms = sum(ms_cls[i] for i in range(7))
    for k, v in kern.items():
        v["share"] = v["ms_total"] / total_kernel_ms if total_kernel_ms else None
    dominant = max(("net_recurrent", "backup_select"), key=lambda k: kern[k]["ms_total"])
    peaks = measured_peaks()

    def roofline_of(name):
        per_launch_s = kern[name]["us_per_launch"] * 1e-6
        if name == "net_recurrent":
            achieved = FLOP_PER_SIM * B / per_launch_s / 1e12
            peak = peaks["bf16_tflops_sustained"]
            return {"kernel": "net_tc<recurrent> (fused g + reward/policy/value heads, tcgen05)", "bound": "tensor",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": peaks["source"] + ", sustained bf16",
                    "algorithmic": f"{FLOP_PER_SIM} FLOP/sim x {B} sims per launch"}
        # fused tree kernel, algorithmic bytes per simulation: selection reads one 128 B record per
        # level, backup reads + writes one 16 B slot per level, the expansion writes one 128 B record, plus the
        # per-search scalars (leaf ids 5 B r/w, p 24 B, r/v 8 B, min/max + root W 24 B r/w, path 4 B/level r/w)
        depth = 3.4
        bytes_per_sim = 128 * depth + 32 * depth + 128 + 8 * depth + 10 + 32 + 48
        achieved = bytes_per_sim * B / per_launch_s / 1e9
        return {"kernel": "search_backup_select (expand + backup + next select)", "bound": "hbm", "achieved": achieved,
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": None,
                "peak_source": peaks["source"],
                "algorithmic": f"{bytes_per_sim:.0f} B/sim x {B} sims per launch (mean leaf depth {depth})"}

    try:  # DRAM traffic per launch from the committed ncu --set full capture of this workload (profiles/)
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            traffic = json.load(f)
    except (OSError, ValueError):
        traffic = {}

    def with_traffic(r, name):
        t = traffic.get(name, {}).get("dram_bytes_per_launch") if B == GAMES_PER_GPU and S == N_SIMS else None
        r["traffic"] = t
        r["traffic_source"] = traffic.get("source") if t else None
        return r

    roofline = with_traffic(roofline_of(dominant), dominant)
    other = "backup_select" if dominant == "net_recurrent" else "net_recurrent"
    roofline_other = with_traffic(roofline_of(other), other)
    # ---- end to end: env words from pinned host memory in, move records back to the host, every step
    h_words = torch.empty(B, dtype=torch.int32).pin_memory()
    h_words.copy_(sp.env.words.cpu())
    h_out = {k: torch.empty_like(v, device="cpu").pin_memory() for k, v in sp.slot(0).items()}
    h_next = torch.empty(B, dtype=torch.int32).pin_memory()

    def e2e_step():
        sp.env.words.copy_(h_words, non_blocking=True)
        t = sp.move()
        for k, v in sp.slot(t).items():
            h_out[k].copy_(v, non_blocking=True)
        h_next.copy_(sp.env.words, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        h_words.copy_(h_next)  # the host owns the env state between steps
        if world > 1:
            hdist.all_gather_records(sp.slot(t), gather_buf)

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_rate = world * B * S * args.steps / (ms_e2e * 1e-3)
    d2h = sum(v.numel() * v.element_size() for v in h_out.values()) + 4 * B

    # ---- raw env throughput (BASELINE.json configs[3]: N=10, 2^24 envs, random legal moves)
    env_line = None
    if not args.no_env:
        nenv = 1 << 24
        env = VecHanoi(10, 200, nenv, dev)
        env.reset()
        acts = torch.randint(0, 6, (nenv,), dtype=torch.uint8, device=dev)
        k_env = [0]

        def env_step_fixed():
            env.step(acts, want_obs=False)

        def env_step_rand():
            env.step_random(seed=1, step_index=k_env[0])
            k_env[0] += 1

        for f in (env_step_fixed, env_step_rand):
            for _ in range(3):
                f()
        ms_env = timed(env_step_fixed, 20) / 20
        ms_rand = timed(env_step_rand, 20) / 20
        ms_roll = timed(lambda: env.rollout_random(64, seed=1, step_index=0), 3) / 3
        gbs = ENV_BYTES_PER_STEP * nenv / (ms_env * 1e-3) / 1e9
        env_line = {"workload": "hanoi10_2^24envs (BASELINE.json configs[3])", "n_envs_per_gpu": nenv,
                    "step_given_actions_per_s": world * nenv / (ms_env * 1e-3),
                    "step_random_legal_per_s": world * nenv / (ms_rand * 1e-3),
                    "fused_rollout64_random_legal_per_s": world * nenv * 64 / (ms_roll * 1e-3),
                    "roofline": {"kernel": "env_step_vec4", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                                 "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                                 "algorithmic": f"{ENV_BYTES_PER_STEP} B/step x {nenv} steps per launch"}}
        del env, acts

    # ---- §8f rows (episode post-processing, replay ring, acting harness): measured only on request
    extras = None
    if args.extras and rank == 0:
        extras = measure_extras(dev, weights, peaks, timed)

    if rank == 0:
        line = {
            "metric": METRIC, "value": sims_per_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.mode == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks,
            "e2e": {"value": e2e_rate, "unit": UNIT, "h2d_bytes_per_step": 4 * B, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "roofline": roofline, "roofline_second_kernel": roofline_other, "kernels": kern,
            "env_steps_per_second": world * B * args.steps / (ms * 1e-3), "env": env_line,
            "targets": {"sims_per_s_8gpu": 1e8, "env_steps_per_s_8gpu": 1e9},
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        if extras is not None:
            line["extras"] = extras
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--mode", choices=["bf16", "fp32"], default=os.environ.get("HMZ_BENCH_MODE", "bf16"))
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU, help="games per GPU (default: BASELINE config)")
    ap.add_argument("--sims", type=int, default=N_SIMS)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-env", action="store_true")
    ap.add_argument("--extras", action="store_true", help="also time the SURVEY §8f rows (episode kernels, replay ring, acting harness)")
    ap.add_argument("--groups", type=int, default=0, help="concurrent search groups in hmz_search_run (0 = auto)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
