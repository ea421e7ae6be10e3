"""Tensor-core kernel regression check: outputs on random rows vs a saved reference (bit for bit) and launch time.

    python tools/tc_compare.py save     # writes gpurun_out/tc_ref.npz with the current build; copy it to tools/_tc_ref.npz
    python tools/tc_compare.py check    # compares the current build with tools/_tc_ref.npz
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from muzero_hanoi_b200.engine import PackedWeights, VecHanoi
from oracle import port

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "tc_ref.npz")  # written on the GPU box
REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_tc_ref.npz")  # copy it here (git-ignored): gpurun_out/ does not travel


def run(n=int(os.environ.get("N_ROWS", 9000))):
    torch.manual_seed(0)
    w = PackedWeights(port.make_weights(5, 3), 5, 1)
    E = 4
    lat = torch.rand(n, E, 64, device="cuda").to(torch.bfloat16)
    rows = torch.randint(0, E - 1, (n,), dtype=torch.int16, device="cuda")
    acts = torch.randint(0, 6, (n,), dtype=torch.uint8, device="cuda")
    r, v, p = torch.empty(n, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, 6, device="cuda")

    def rec():
        w.recurrent(n, latents_in=lat, in_rows_per_item=E, in_row=rows, actions=acts, latents_out=lat, out_rows_per_item=E,
                    out_row=E - 1, latent_dtype=1, r=r, p=p, v=v)
    for _ in range(3):
        rec()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        rec()
    e1.record()
    torch.cuda.synchronize()
    print("recurrent us/launch %.2f" % (e0.elapsed_time(e1) * 1000 / 20), flush=True)
    out = dict(h=lat[:, E - 1].float().cpu().numpy(), r=r.cpu().numpy(), v=v.cpu().numpy(), p=p.cpu().numpy())
    env = VecHanoi(5, 200, n)
    env.random_reset(seed=1)
    h0 = torch.zeros(n, 2, 64, device="cuda", dtype=torch.bfloat16)
    p0, v0 = torch.empty(n, 6, device="cuda"), torch.empty(n, device="cuda")
    w.initial(n, words=env.words, latents_out=h0, out_rows_per_item=2, latent_dtype=1, p0=p0, v0=v0)
    torch.cuda.synchronize()
    out.update(h0=h0.float().cpu().numpy(), p0=p0.cpu().numpy(), v0=v0.cpu().numpy())
    m = 1000  # ragged size: the last tile pair is partly empty
    r2, v2, p2 = torch.empty(m, device="cuda"), torch.empty(m, device="cuda"), torch.empty(m, 6, device="cuda")
    lat2 = lat[:m].clone()
    w.recurrent(m, latents_in=lat2, in_rows_per_item=E, in_row=rows[:m].contiguous(), actions=acts[:m].contiguous(), latents_out=lat2,
                out_rows_per_item=E, out_row=E - 1, latent_dtype=1, r=r2, p=p2, v=v2)
    torch.cuda.synchronize()
    out.update(h_r=lat2[:, E - 1].float().cpu().numpy(), r_r=r2.cpu().numpy(), v_r=v2.cpu().numpy(), p_r=p2.cpu().numpy())
    return out


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    out = run()
    if mode == "save":
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        np.savez_compressed(OUT, **out)
        print("saved", OUT)
    else:
        ref = np.load(REF)
        bad = 0
        for k in out:
            d = np.abs(out[k] - ref[k])
            print(k, "max |new - ref| = %.3e" % d.max(), "mismatching = %d / %d" % ((d > 0).sum(), d.size))
            bad += int((d > 0).sum())
        print("BIT-IDENTICAL" if bad == 0 else "DIFFERENT")
