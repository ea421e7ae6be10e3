"""Phase timeline (clock64 deltas) of one lane pair of the fused backup + select kernel over several simulations."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from muzero_hanoi_b200 import _lib
from muzero_hanoi_b200.engine import BatchedMCTS, PackedWeights, VecHanoi
from oracle import port

B, S, n = int(os.environ.get("B", 65536)), 100, 5
lib = _lib.load()
if os.environ.get("TORCH_INIT"):  # the bench's network: torch's default initialisation (shallower trees)
    from muzero_hanoi_b200.networks import MuZeroNet
    torch.manual_seed(0)
    w = PackedWeights(MuZeroNet(3 * n, 6, 0.002, "cpu", TD_return=True).state_dict(), n, 1)
else:
    w = PackedWeights(port.make_weights(n, 3), n, 1)
env = VecHanoi(n, 200, B)
env.reset()
env.random_reset(seed=5)
m = BatchedMCTS(0.8, 0.25, S, B, latent_dtype=1)
m.store.set_schedule(int(os.environ.get("GROUPS", 1)))
noise = torch.from_numpy(np.random.default_rng(0).dirichlet(np.full(6, 0.25), B)).cuda()
uni = torch.rand(B, dtype=torch.float64, device="cuda")
run = lambda: m.run_mcts(w, words=env.words, temperature=1.0, deterministic=False, noise=noise, uniforms=uni)
run()
torch.cuda.synchronize()
SIM = int(os.environ.get("SIM", 80))
buf = (C.c_ulonglong * 64)()
for search in [int(x) for x in os.environ.get("SEARCHES", "0,1,17,5000,30000,30001,65535").split(",")]:
    _lib.check(lib.hmz_debug_tree_timeline(search | (SIM << 32), None))
    run()
    torch.cuda.synchronize()
    _lib.check(lib.hmz_debug_tree_timeline(-1, buf))
    t = np.array(list(buf), dtype=np.int64)
    print(f"search {search}: backed-up depth {t[2]}, next path depth {t[4]}")
    t0 = t[0]
    marks = [(0, "entry"), (1, "leaf scalars loaded"), (6, "fresh record written, r loaded"), (7, "root W / min-max loaded"),
             (3, "backup done (stores issued)"), (40, "walk entered (min / max exchanged)"), (41, "level 0 loads issued")]
    for bt in range(4):
        marks += [(24 + 3 * bt, f"backup batch {bt}: path entries loaded"), (25 + 3 * bt, f"backup batch {bt}: slots loaded"),
                  (26 + 3 * bt, f"backup batch {bt}: recurrence done")]
    for d in range(8):
        marks += [(8 + 2 * d, f"select level {d}: record loaded"), (9 + 2 * d, f"select level {d}: arg-max done")]
    marks += [(5, "end")]
    prev = 0
    for k, nm in sorted(marks, key=lambda kn: t[kn[0]] if t[kn[0]] >= t0 else 1 << 62):
        if t[k] < t0:
            continue
        print(f"  {t[k] - t0:7d} (+{t[k] - t0 - prev:6d})  {nm}")
        prev = t[k] - t0
