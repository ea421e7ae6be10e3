"""Per-launch time of recurrent_inference in the three network modes (FFMA float32, tcgen05 three-part float32,
tcgen05 bf16) over batch sizes: python tools/x3_probe.py  (GPU box; CUDA events around 20 launches after 3 warm-ups)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_hanoi_b200 import _lib  # noqa: E402
from muzero_hanoi_b200.engine import PackedWeights  # noqa: E402
from muzero_hanoi_b200.networks import MuZeroNet  # noqa: E402

FLOP = 203776
torch.manual_seed(0)
sd = MuZeroNet(15, 6, 0.002, "cpu", TD_return=True).state_dict()
dev = torch.device("cuda")
ROWS = [int(x) for x in os.environ["ROWS"].split(",")] if os.environ.get("ROWS") else [1024, 4096, 8192, 16384, 18944, 32768, 65536, 131072]
MODES = os.environ.get("MODES", "ffma,x3,bf16").split(",")
for rows in ROWS:
    g = torch.Generator().manual_seed(1)
    h_in = torch.rand(rows, 64, generator=g).to(dev)
    acts = torch.randint(0, 6, (rows,), generator=g).to(torch.uint8).to(dev)
    line = [f"rows {rows:7d}"]
    for name, md, ld in (("ffma", _lib.MODE_FP32, 0), ("x3", _lib.MODE_FP32X3, 0), ("bf16", _lib.MODE_BF16, 1)):
        if name not in MODES:
            continue
        w = PackedWeights(sd, 5, md, dev)
        src = h_in.to(torch.bfloat16) if ld else h_in
        h = torch.empty(rows, 64, device=dev, dtype=torch.bfloat16 if ld else torch.float32)
        r, v, p = torch.empty(rows, device=dev), torch.empty(rows, device=dev), torch.empty(rows, 6, device=dev)
        run = lambda: w.recurrent(rows, latents_in=src, in_rows_per_item=1, in_row=None, actions=acts, latents_out=h,
                                  out_rows_per_item=1, out_row=0, latent_dtype=ld, r=r, p=p, v=v)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        line.append(f"{name} {us:8.1f} us ({FLOP * rows / us / 1e6:6.1f} TFLOP/s eff.)")
    print("  ".join(line), flush=True)
