"""Prints the phase timeline (clock64 deltas, CTA 0) of one net_recurrent_tc launch.  HMZ_TC_TIMELINE=1."""
import ctypes as C, os, sys
os.environ["HMZ_TC_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from muzero_hanoi_b200 import _lib
from muzero_hanoi_b200.engine import PackedWeights
from muzero_hanoi_b200.networks import MuZeroNet
torch.manual_seed(0)
n = 65536
net = MuZeroNet(15, 6, 0.002, "cpu", TD_return=True)
w = PackedWeights(net.state_dict(), 5, 1)
h_in = torch.rand(n, 64, device="cuda").to(torch.bfloat16)
acts = torch.randint(0, 6, (n,), dtype=torch.uint8, device="cuda")
h = torch.empty(n, 64, device="cuda", dtype=torch.bfloat16)
r, v, p = torch.empty(n, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, 6, device="cuda")
for _ in range(3):
    w.recurrent(n, latents_in=h_in, in_rows_per_item=1, in_row=None, actions=acts, latents_out=h, out_rows_per_item=1,
                out_row=0, latent_dtype=1, r=r, p=p, v=v)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 96)()
_lib.check(_lib.load().hmz_debug_tc_timeline(buf))
t = np.array(list(buf), dtype=np.int64)
names = {0: "ctl start", 1: "ctl A0 gathered", 2: "ctl Wg1 landed", 3: "ctl L1 issued", 4: "ctl L1 complete", 5: "ctl Wg2 landed",
         6: "ctl g-hid h0 written", 7: "ctl g-hid h1 written", 8: "ctl L2 complete", 9: "ctl E2 done"}
for hd, nm in enumerate("rpv"):
    for j, what in enumerate(["W1 landed", "first MMA complete", "W2 landed", "hid h0 written", "hid h1 written", "second MMA complete"]):
        names[10 + hd * 6 + j] = f"ctl {nm}: {what}"
names.update({32: "epi gather done", 33: "epi saw L1", 34: "epi g-hid math done", 35: "epi fenced+arrived", 36: "epi saw L2", 37: "epi E2 done", 63: "end"})
for hd, nm in enumerate("rpv"):
    for j, what in enumerate(["saw first layer", "hidden epilogue done", "saw second layer", "final epilogue done"]):
        names[38 + hd * 4 + j] = f"epi {nm}: {what}"
t0 = t[0]
ev = sorted((int(t[k] - t0), names[k]) for k in names if t[k] != 0)
prev = 0
for c, nm in ev:
    print(f"{c:8d}  (+{c - prev:6d})  {nm}")
    prev = c
