"""One lambertian SDS step between cudaProfilerStart/Stop, for `ncu --profile-from-start off ...` (launch list / full captures).
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stable-dreamfusion_b200")]
import torch
from guidance.sd_utils import StableDiffusion
from sdf_b200.options import default_opt
from sdf_b200.trainer import SDSTrainer

dev = torch.device("cuda:0")
shading = sys.argv[1] if len(sys.argv) > 1 else "lambertian"
dmtet = len(sys.argv) > 2 and sys.argv[2] == "dmtet"              # python tools/profile_step.py lambertian dmtet  -> one C5 step
if dmtet:
    from sdf_b200.options import dmtet_opt
    guidance = StableDiffusion(dev, weights="random", n_views=1, render_hw=512, seed=0, capture=False)
    tr = SDSTrainer(dmtet_opt(), dev, guidance, seed=0)
    tr.model.update_extra_state()
    tr.model.init_tet()
else:
    guidance = StableDiffusion(dev, weights="random", n_views=1, render_hw=64, seed=0, capture=False)     # eager lists: every launch is visible
    tr = SDSTrainer(default_opt(h=64, w=64), dev, guidance, seed=0)
for s in [shading if dmtet else "latent", shading, shading]:
    tr.train_step(shading=s)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.train_step(shading=shading)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one", shading, "step; M =", tr.last_M)
