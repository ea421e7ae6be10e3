"""Phase timeline (clock64, CTA 0, first tile) of one net_x3_recurrent launch (HMZ_MODE_FP32X3): python tools/x3_timeline.py
Columns per chunk g: first-layer issue from / to, hidden epilogue from / to, second-layer issue from / to (clk after kernel entry)."""
import ctypes as C
import os
import sys

os.environ["HMZ_X3_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from muzero_hanoi_b200 import _lib  # noqa: E402
from muzero_hanoi_b200.engine import PackedWeights  # noqa: E402
from muzero_hanoi_b200.networks import MuZeroNet  # noqa: E402

n = int(os.environ.get("N", 65536))
torch.manual_seed(0)
w = PackedWeights(MuZeroNet(15, 6, 0.002, "cpu", TD_return=True).state_dict(), 5, _lib.MODE_FP32X3)
h_in = torch.rand(n, 64, device="cuda")
acts = torch.randint(0, 6, (n,), dtype=torch.uint8, device="cuda")
h = torch.empty(n, 64, device="cuda")
r, v, p = torch.empty(n, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, 6, device="cuda")
for _ in range(3):
    w.recurrent(n, latents_in=h_in, in_rows_per_item=1, in_row=None, actions=acts, latents_out=h, out_rows_per_item=1, out_row=0,
                latent_dtype=0, r=r, p=p, v=v)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 160)()
_lib.check(_lib.load().hmz_debug_x3_timeline(buf))
m = np.array(list(buf), dtype=np.int64)
t0 = m[110]
rel = lambda i: int(m[i] - t0)
print(f"rows {n}: prologue done {rel(111)}, gather {rel(104)} -> {rel(105)}")
print(" g net c | W1 TMA issued, landed | L1 loop top, issue from   to |  epilogue from    to |  L2 issue from    to")
for g in range(16):
    print(f"{g:2d}  {'grvp'[g >> 2]}  {g & 3} | {rel(128 + g):8d} {rel(144 + g):8d} | {rel(112 + g):8d} {rel(g):8d} {rel(16 + g):8d} | {rel(64 + g):8d} {rel(80 + g):8d} | {rel(32 + g):8d} {rel(48 + g):8d}")
    if (g & 3) == 3:
        print(f"      output epilogue of network {'grvp'[g >> 2]}: {rel(96 + (g >> 2))} -> {rel(100 + (g >> 2))}"
              + (f" (raw latent tile published {rel(106)})" if g == 3 else ""))
