"""GPU busy share of the bench's training step: the sum of kernel durations (CUPTI, via torch.profiler) over 10 steps of the bench cycle
against the CUDA-event time of the same steps.  A busy share well below 1 means launch gaps (host-bound eager launches), not kernel time.
    python tools/step_busy.py [lambertian]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stable-dreamfusion_b200")]
import torch
from torch.profiler import ProfilerActivity, profile
from guidance.sd_utils import StableDiffusion
from sdf_b200.options import default_opt
from sdf_b200.trainer import SDSTrainer

dev = torch.device("cuda:0")
guidance = StableDiffusion(dev, weights="random", n_views=1, render_hw=64, seed=0, capture=True)
tr = SDSTrainer(default_opt(h=64, w=64), dev, guidance, seed=0)
cycle = ["latent"] + ["lambertian", "textureless", "albedo", "lambertian"] if len(sys.argv) < 2 else [sys.argv[1]]
for i in range(6):
    tr.train_step(shading=cycle[i % len(cycle)], read_loss=False)
torch.cuda.synchronize()
N = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(N):
    tr.train_step(shading=cycle[i % len(cycle)], read_loss=False)
e1.record()
torch.cuda.synchronize()
ms_plain = e0.elapsed_time(e1) / N
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    e0.record()
    for i in range(N):
        tr.train_step(shading=cycle[i % len(cycle)], read_loss=False)
    e1.record()
    torch.cuda.synchronize()
ms_prof = e0.elapsed_time(e1) / N
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
busy = sum(e.device_time for e in evs if hasattr(e, "device_time")) / 1e3 / N
by = {}
for e in evs:
    k = e.name[:60]
    d = by.setdefault(k, [0, 0.0])
    d[0] += 1
    d[1] += e.device_time / 1e3 / N
print(f"ms/step plain {ms_plain:.3f}   under profiler {ms_prof:.3f}   sum of kernel+memcpy durations {busy:.3f} ms/step  -> busy share {busy / ms_plain:.3f}")
for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {t:8.3f} ms/step  {n / N:7.1f} launches/step  {k}")
