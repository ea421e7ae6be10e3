"""Launch Gantt of the group schedule (hmz_debug_gantt): for a window of simulations of one self-play move, every launch of
the network kernel and of the fused tree kernel with its first block's start and last block's end (globaltimer, us), plus
per-kernel busy intervals, how much of the window has a network / a tree kernel in flight, and the gaps.
    B=65536 S=100 SCHEDULE=0 SIMS=40:44 python tools/gantt.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from muzero_hanoi_b200 import _lib
from muzero_hanoi_b200.engine import PackedWeights, SelfPlay
from muzero_hanoi_b200.networks import MuZeroNet

B, S, n = int(os.environ.get("B", 65536)), int(os.environ.get("S", 100)), int(os.environ.get("N", 5))
lo, hi = [int(x) for x in os.environ.get("SIMS", "40:44").split(":")]
lib = _lib.load()
torch.manual_seed(0)
w = PackedWeights(MuZeroNet(3 * n, 6, 0.002, "cpu", TD_return=True).state_dict(), n, _lib.MODE_BF16)
sp = SelfPlay(n, 200, B, S, w, seed=1, ring_slots=4, latent_dtype=_lib.LATENT_BF16)
sp.mcts.store.set_schedule(int(os.environ.get("SCHEDULE", 0)))
for _ in range(3):
    sp.move()
torch.cuda.synchronize()
_lib.check(lib.hmz_debug_gantt(1, None, 0, None))
sp.move()
torch.cuda.synchronize()
cap = 4096
buf = (C.c_ulonglong * (4 * cap))()
n_out = C.c_int(0)
_lib.check(lib.hmz_debug_gantt(0, buf, cap, C.byref(n_out)))
rec = np.array(list(buf[: 4 * n_out.value]), dtype=np.int64).reshape(-1, 4)
rec = rec[rec[:, 2] < (1 << 62)]  # launches that never ran a block keep the reset pattern
kind, tag, t0, t1 = rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3]
sim, grp = tag >> 8, tag & 0xFF
move_span = (t1.max() - t0.min()) / 1e3
print(f"B={B} S={S}: {len(rec)} launches recorded, move search phase {move_span:.1f} us = {move_span / S:.2f} us per simulation round")
sel = (sim >= lo) & (sim < hi)
base = t0[sel].min()
print(f"\nlaunches of simulations [{lo}, {hi}) (us from the first start):")
print("  kind  group sim   start     end   duration")
for i in np.argsort(t0):
    if sel[i]:
        print(f"  {'net ' if kind[i] == 0 else 'tree'}  {grp[i]:3d}  {sim[i]:3d} {(t0[i] - base) / 1e3:8.2f} {(t1[i] - base) / 1e3:8.2f} {(t1[i] - t0[i]) / 1e3:8.2f}")


def union(mask):
    iv = sorted(zip(t0[mask], t1[mask]))
    tot, cur_s, cur_e = 0, None, None
    for s_, e_ in iv:
        if cur_e is None or s_ > cur_e:
            if cur_e is not None:
                tot += cur_e - cur_s
            cur_s, cur_e = s_, e_
        else:
            cur_e = max(cur_e, e_)
    if cur_e is not None:
        tot += cur_e - cur_s
    return tot


steady = (sim >= 10) & (sim < S - 5)
span = t1[steady].max() - t0[steady].min()
print(f"\nsteady state (simulations 10..{S - 6}): {span / 1e3 / (S - 15):.2f} us per round")
for k, nm in ((0, "network"), (1, "tree")):
    m = steady & (kind == k)
    if not m.any():  # (server / persistent schedules: the network CTAs are one resident launch, not recorded)
        continue
    d = (t1[m] - t0[m]) / 1e3
    print(f"  {nm:8s}: {m.sum()} launches, duration mean {d.mean():.2f} us (min {d.min():.2f}, max {d.max():.2f}); "
          f"some {nm} kernel in flight {100 * union(m) / span:.1f} % of the time; sum of durations / span = {d.sum() * 1e3 / span:.2f} kernels in flight on average")
both = union(steady)
print(f"  any kernel in flight {100 * both / span:.1f} % of the time")
# per group: gap between a tree kernel's end and the next network kernel's start (and the reverse)
for g in sorted(set(grp[steady])):
    mg = steady & (grp == g)
    order = np.argsort(t0[mg])
    k_, s0, s1 = kind[mg][order], t0[mg][order], t1[mg][order]
    gaps_tn = [(s0[i + 1] - s1[i]) / 1e3 for i in range(len(k_) - 1) if k_[i] == 1 and k_[i + 1] == 0]
    gaps_nt = [(s0[i + 1] - s1[i]) / 1e3 for i in range(len(k_) - 1) if k_[i] == 0 and k_[i + 1] == 1]
    if not gaps_tn:
        gaps_tt = [(s0[i + 1] - s1[i]) / 1e3 for i in range(len(k_) - 1)]
        print(f"  group {g}: tree launch end -> next tree launch's first block {np.mean(gaps_tt):6.2f} us")
        continue
    print(f"  group {g}: tree end -> next network start {np.mean(gaps_tn):6.2f} us (first block), network end -> tree start {np.mean(gaps_nt):6.2f} us "
          f"(negative = the dependent was already resident, waiting)")
