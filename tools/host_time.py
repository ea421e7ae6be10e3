"""Host-side cost of one SelfPlay.move() (launch-bound check): wall time of the call without synchronising vs GPU time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from muzero_hanoi_b200 import _lib
from muzero_hanoi_b200.engine import PackedWeights, SelfPlay
from oracle import port
B, S, n = 65536, 100, 5
w = PackedWeights(port.make_weights(n, 3), n, 1)
for groups in (1, 4, _lib.SCHEDULE_PERSISTENT):
    sp = SelfPlay(n, 200, B, S, w, seed=1, latent_dtype=1)
    sp.mcts.store.set_schedule(groups)
    for _ in range(3):
        sp.move()
    torch.cuda.synchronize()
    host, gpu = [], []
    n0 = _lib.load().hmz_launch_count()
    for _ in range(5):
        t0 = time.perf_counter()
        sp.move()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        host.append((t1 - t0) * 1e3)
        gpu.append((t2 - t0) * 1e3)
    print(f"groups={groups}: host enqueue {min(host):.2f} ms, until GPU done {min(gpu):.2f} ms, launches/move {(_lib.load().hmz_launch_count() - n0) // 5}")
