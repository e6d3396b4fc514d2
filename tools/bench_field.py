"""Fused field kernels alone: forward / backward time on a typical sample set (CUDA events, 5 reps, median).
    python tools/bench_field.py [M] [shading]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stable-dreamfusion_b200")]
import torch
from sdf_b200 import field
from sdf_b200.ngp import InstantNGP
from sdf_b200.options import default_opt

M = int(sys.argv[1]) if len(sys.argv) > 1 else 432000
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = InstantNGP(default_opt(h=64, w=64)).to(dev)
xyz = (torch.rand(M, 3, device=dev) * 2 - 1) * 0.5
l = torch.nn.functional.normalize(torch.randn(M, 3, device=dev), dim=-1)
for shading in (sys.argv[2:] or ["albedo", "lambertian", "normal"]):
    tf, tb = [], []
    for rep in range(7):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        s, c, n = m(xyz, None, l, ratio=0.5, shading=shading)
        e[1].record()
        loss = s.sum() + c.sum() + (n.sum() if n is not None else 0.0)
        loss.backward()
        e[2].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
    tf.sort(); tb.sort()
    np_ = 1 if shading == "albedo" else 7
    f, b = tf[len(tf) // 2], tb[len(tb) // 2]
    print(f"{shading:12s} M={M} fwd {f:7.3f} ms ({M * np_ * 540 / f / 1e6:7.1f} GB/s alg)  bwd {b:7.3f} ms ({M * np_ * 1052 / b / 1e6:7.1f} GB/s alg)  [bwd includes the loss reductions and autograd wrapper]", flush=True)
