import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from muzero_hanoi_b200 import _lib
from muzero_hanoi_b200.engine import BatchedMCTS, PackedWeights, VecHanoi
from muzero_hanoi_b200.networks import MuZeroNet
torch.manual_seed(0)
n, B, S = 5, 65536, 100
net = MuZeroNet(3 * n, 6, 0.002, "cpu", TD_return=True)
w = PackedWeights(net.state_dict(), n, 1)
env = VecHanoi(n, 200, B); env.reset()
m = BatchedMCTS(0.8, 0.25, S, B, latent_dtype=1)
st = m.store
import ctypes as C
p0, v0 = torch.empty(B, 6, device="cuda"), torch.empty(B, device="cuda")
w.initial(B, words=env.words, latents_out=st.latents, out_rows_per_item=st.n_records, latent_dtype=1, p0=p0, v0=v0)
noise = torch.from_numpy(np.random.default_rng(0).dirichlet(np.full(6, .25), size=B)).cuda()
st.desc.root_prior_is_f64 = 1
_lib.check(m.lib.hmz_search_begin_p0(C.byref(st.desc), _lib.ptr(p0), _lib.ptr(noise), 0.25, _lib.current_stream()))
r, v, p = torch.empty(B, device="cuda"), torch.empty(B, device="cuda"), torch.empty(B, 6, device="cuda")
for s in range(S):
    m.select(s)
    if s % 10 == 9 or s == S - 1:
        d = m.leaf_depth.cpu().numpy().astype(np.int64)
        print(f"sim {s}: mean depth {d.mean():.2f}  p99 {np.percentile(d,99):.0f}  max {d.max()}  >8: {(d>8).mean():.4f} >16: {(d>16).mean():.5f} >32: {(d>32).sum()}")
    w.recurrent(B, latents_in=st.latents, in_rows_per_item=st.n_records, in_row=m.leaf_parent, actions=m.leaf_action,
                latents_out=st.latents, out_rows_per_item=st.n_records, out_row=s + 1, latent_dtype=1, r=r, p=p, v=v)
    m.expand_backup(s, r, p, v)
