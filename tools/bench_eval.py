"""Inference render rate of the instant-NGP backbone (the reference's only published number is
"~10 FPS at 800x800", readme.md:28, V100): march / field / composite loop of nerf/renderer.py:759-794 with device-side bookkeeping
(sdf_b200/render_eval.py: no host sync per iteration, on-device alive-ray compaction), random camera on the unit-sphere blob scene after a few occupancy refreshes."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stable-dreamfusion_b200")]
import numpy as np
import torch
from sdf_b200 import synth
from sdf_b200.ngp import InstantNGP
from sdf_b200.options import default_opt

H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 800
dev = torch.device("cuda:0")
torch.manual_seed(0)
opt = default_opt(h=64, w=64)
m = InstantNGP(opt).to(dev)
with torch.no_grad():
    m.encoder.embeddings.uniform_(-0.5, 0.5)          # "trained-like" table so that all levels matter
m.invalidate_mirror()
m.train()
for _ in range(3):
    m.update_extra_state()
m.eval()
rng = np.random.default_rng(0)
times = []
for shading in ("albedo", "lambertian"):
    for rep in range(6):
        pose, _ = synth.rand_pose(rng)
        ro, rd = synth.get_rays(pose, H, W, 20.0)
        ro_t, rd_t = torch.from_numpy(ro).to(dev)[None], torch.from_numpy(rd).to(dev)[None]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            light = torch.nn.functional.normalize(ro_t[0, 0] + torch.randn(3, device=dev), dim=-1)          # one light per frame (nerf/utils.py test_step)
            out = m.render(ro_t, rd_t, None, H, W, staged=False, perturb=False, bg_color=None, ambient_ratio=0.5, shading=shading, light_d=light)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rep >= 2:
            times.append((shading, dt))
    ts = [t for s, t in times if s == shading]
    print(f"{H}x{W} {shading:10s}: {1.0 / (sum(ts) / len(ts)):7.1f} FPS ({1e3 * sum(ts) / len(ts):.1f} ms/frame, coverage {float((out['weights_sum'] > 0.5).float().mean()):.2f})", flush=True)
