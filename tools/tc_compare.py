"""v3 vs v4 tensor-core kernel: output agreement on random rows and launch time (one process per variant)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import numpy as np, torch
    from muzero_hanoi_b200.engine import PackedWeights
    from oracle import port
    n = int(os.environ.get("N_ROWS", 65536))
    torch.manual_seed(0)
    w = PackedWeights(port.make_weights(5, 3), 5, 1)
    E = 4
    lat = torch.rand(n, E, 64, device="cuda").to(torch.bfloat16)
    rows = torch.randint(0, E - 1, (n,), dtype=torch.int16, device="cuda")
    acts = torch.randint(0, 6, (n,), dtype=torch.uint8, device="cuda")
    r, v, p = torch.empty(n, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, 6, device="cuda")

    def run():
        w.recurrent(n, latents_in=lat, in_rows_per_item=E, in_row=rows, actions=acts, latents_out=lat, out_rows_per_item=E,
                    out_row=E - 1, latent_dtype=1, r=r, p=p, v=v)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    print("us/launch %.2f" % (e0.elapsed_time(e1) * 1000 / 20), flush=True)
    np.savez(os.environ["OUT"], h=lat[:, E - 1].float().cpu().numpy(), r=r.cpu().numpy(), v=v.cpu().numpy(), p=p.cpu().numpy())


if __name__ == "__main__":
    if os.environ.get("CHILD"):
        child()
        sys.exit(0)
    import numpy as np
    for v3 in ("1", "0"):
        env = dict(os.environ, CHILD="1", HMZ_TC_V3=v3, OUT=f"/tmp/tc_v3_{v3}.npz")
        print("HMZ_TC_V3=" + v3, flush=True)
        subprocess.run([sys.executable, __file__], env=env, check=True, timeout=300)
    a, b = np.load("/tmp/tc_v3_1.npz"), np.load("/tmp/tc_v3_0.npz")
    for k in ("h", "r", "v", "p"):
        d = np.abs(a[k] - b[k])
        print(k, "max |v3 - v4| = %.3e" % d.max(), "mismatching = %d / %d" % ((d > 0).sum(), d.size))
