"""A/B timing of the three SD launch lists as the bench replays them (CUDA graphs): UNet forward, VAE forward, VAE data-gradient.
Every variant runs in its own process (the kernels read their switches once):
    python tools/bench_lists.py                      # default switches vs SDF_STRIDED_TMA_CONV=0 vs SDF_FLASH_TC_V=1
    python tools/bench_lists.py one                  # this process's environment only
Median of 30 replays per list, CUDA events, 256 MB L2 flush between replays."""
import os
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stable-dreamfusion_b200")]

VARIANTS = [{}, {"SDF_GN_WAVES": "1"}, {"SDF_GN_WAVES": "3"}]
if len(sys.argv) == 1:
    for v in VARIANTS:
        subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=dict(os.environ, **v, BENCH_LISTS_TAG=repr(v)), timeout=600)
    sys.exit(0)

import torch
from guidance.sd_utils import StableDiffusion
dev = torch.device("cuda:0")
g = StableDiffusion(dev, weights="random", n_views=1, render_hw=64, seed=0, capture=True)
eng = g.engine
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = []
for lname, rl in (("unet", eng.unet.runlist), ("vae_fwd", eng.vae.fwd), ("vae_bwd", eng.vae.bwd)):
    for _ in range(3):
        rl.run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rl.run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    out.append(f"{lname} {ts[len(ts) // 2]:.3f} ms (min {ts[0]:.3f})")
print(os.environ.get("BENCH_LISTS_TAG", "{}"), " | ".join(out), flush=True)
