"""Fused field forward / backward on REAL sample sets (marched through the step-0 blob occupancy from a few cameras) vs uniformly random points
of the same count: the real samples cluster inside a small ball, so the backward's scattered reductions pile up on the coarse levels' few
entries.  CUDA events around the bare C-ABI launches, median of 7.
    python tools/bench_field_scene.py [hw]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stable-dreamfusion_b200")]
import numpy as np
import torch
import raymarching
from sdf_b200 import _lib, synth
from sdf_b200.ngp import InstantNGP
from sdf_b200.options import default_opt

hw = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = InstantNGP(default_opt(h=hw, w=hw)).to(dev)
m.update_extra_state()
P = _lib.ptr
c = m.field_cfg()
sn = m.sigma_net.net
wts = [P(t) for t in (sn[0].weight, sn[0].bias, sn[1].weight, sn[1].bias, sn[2].weight, sn[2].bias)]
table = m.table_half()
st = _lib.stream()


def run(xyz, tag):
    M = xyz.shape[0]
    g = torch.Generator(device=dev).manual_seed(1)
    l = torch.nn.functional.normalize(torch.randn(M, 3, device=dev, generator=g), dim=-1).contiguous()
    sig, col, nrm, aux = torch.empty(M, device=dev), torch.empty(M, 3, device=dev), torch.empty(M, 3, device=dev), torch.empty(M, 10, device=dev)
    feat = torch.empty(_lib.query("sdf_field_feat_bytes", M, 1) // 4, device=dev, dtype=torch.int32)
    gs, gc = torch.randn(M, device=dev, generator=g) * 1e-3, torch.randn(M, 3, device=dev, generator=g)
    gt = torch.zeros_like(m.encoder.embeddings)
    gw = [torch.zeros_like(t) for t in (sn[0].weight, sn[0].bias, sn[1].weight, sn[1].bias, sn[2].weight, sn[2].bias)]
    fargs = (P(xyz), M, None, P(table), P(c["offsets"]), c["L"], c["levels_active"], c["S"], int(c["H"]), int(c["smoothstep"]), *wts, m.bound, c["blob_density"],
             c["blob_radius"], 1, P(l), 1, 0.5)
    tf, tb = [], []
    for rep in range(9):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        _lib.call("sdf_field_forward", *fargs, P(sig), P(col), P(nrm), P(aux), P(feat), st)
        e[1].record()
        _lib.call("sdf_field_backward", *fargs, P(aux), P(gs), P(gc), None, P(gt), *[P(t) for t in gw], P(feat), st)
        e[2].record()
        torch.cuda.synchronize()
        if rep >= 2:
            tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
    tf.sort(); tb.sort()
    print(f"{tag:34s} M={M:8d}  fwd {tf[len(tf) // 2]:6.3f} ms  bwd {tb[len(tb) // 2]:6.3f} ms  ({M * 7 * 1052 / tb[len(tb) // 2] / 1e6:7.1f} GB/s alg)", flush=True)


for (r, th, ph, fov) in [(3.2, 90.0, 0.0, 20.0), (3.0, 60.0, 120.0, 12.0), (3.5, 100.0, -60.0, 28.0)]:
    ro, rd = synth.get_rays(synth.circle_pose(r, th, ph), hw, hw, fov)
    ro_t, rd_t = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    nears, fars = raymarching.near_far_from_aabb(ro_t, rd_t, m.aabb_train, 0.2)
    xyz, _, _, _ = raymarching.march_rays_train(ro_t, rd_t, 1.0, m.density_bitfield, 1, 128, nears, fars, True, 0, 1024)
    xyz = xyz.contiguous()
    run(xyz, f"marched r={r} th={th} fov={fov}")
    M = xyz.shape[0]
    g = torch.Generator(device=dev).manual_seed(2)
    run(((torch.rand(M, 3, device=dev, generator=g) * 2 - 1) * 0.5).contiguous(), "uniform in [-0.5,0.5]^3")
    run(xyz[torch.randperm(M, device=dev)].contiguous(), "marched, shuffled order")
