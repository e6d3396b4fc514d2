"""Golden vectors for the fused Adan step from the reference's own optimizer (optimizer.py:23-258, foreach=False, as main.py:368
constructs it): run HERE (CPU, needs /root/reference):  python tests/golden/make_golden_adan.py
Writes tests/golden/adan.npz: initial parameters, per-step gradients (one step with a global norm above max_grad_norm so that the
clip is exercised), per-group learning rates, and the REFERENCE's parameters after every step."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("SDF_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)
from optimizer import Adan  # noqa: E402  (the reference's)

g = torch.Generator().manual_seed(7)
shapes = [(3000, 2), (64, 32), (64,), (5,)]
lrs = [1e-2, 1e-3, 1e-3, 1e-3]
params = [torch.nn.Parameter(torch.randn(*s, generator=g) * 0.1) for s in shapes]
opt = Adan([{"params": [p], "lr": lr} for p, lr in zip(params, lrs)], eps=1e-8, weight_decay=2e-5, max_grad_norm=5.0, foreach=False)
out = {"lrs": np.array(lrs, np.float64)}
for i, p in enumerate(params):
    out[f"p0.{i}"] = p.detach().numpy().copy()
for step in range(4):
    scale = 40.0 if step == 1 else 1.0           # step 1: ||g|| >> 5 -> clipped
    for i, p in enumerate(params):
        p.grad = torch.randn(*shapes[i], generator=g) * scale * (0.1 if i == 0 else 1.0)
        out[f"g{step}.{i}"] = p.grad.numpy().copy()
    opt.step()
    for i, p in enumerate(params):
        out[f"p{step + 1}.{i}"] = p.detach().numpy().copy()
path = os.path.join(ROOT, "tests", "golden", "adan.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB")
