"""GPU: fused Adan (csrc/adan.cu via sdf_b200.optimizer.Adan) against a float64 restatement of the reference's
_single_tensor_adan (optimizer.py:201-258) + global-norm clipping (optimizer.py:110-131), over several steps and two groups."""
import math

import pytest
import torch

from sdf_b200.optimizer import Adan

pytestmark = pytest.mark.gpu


def ref_adan_steps(params, grads_per_step, lrs, betas=(0.98, 0.92, 0.99), eps=1e-8, wd=2e-5, max_norm=5.0):
    p = [x.double().clone() for x in params]
    m = [torch.zeros_like(x) for x in p]; d = [torch.zeros_like(x) for x in p]; n = [torch.zeros_like(x) for x in p]
    pre = [None] * len(p)
    b1, b2, b3 = betas
    for step, grads in enumerate(grads_per_step, 1):
        gn = math.sqrt(sum((g.double() ** 2).sum().item() for g in grads))
        clip = min(max_norm / (gn + eps), 1.0)
        bc1, bc2, bc3s = 1 - b1 ** step, 1 - b2 ** step, math.sqrt(1 - b3 ** step)
        for i, g in enumerate(grads):
            g = g.double() * clip
            if pre[i] is None:
                pre[i] = -g
            diff = pre[i] + g
            m[i] = b1 * m[i] + (1 - b1) * g
            d[i] = b2 * d[i] + (1 - b2) * diff
            u = b2 * diff + g
            n[i] = b3 * n[i] + (1 - b3) * u * u
            denom = n[i].sqrt() / bc3s + eps
            p[i] = (p[i] - lrs[i] / bc1 * m[i] / denom - lrs[i] * b2 / bc2 * d[i] / denom) / (1 + lrs[i] * wd)
            pre[i] = -g
    return p


def test_adan_matches_reference_math(device):
    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.randn(100003, 2, device=device) * 1e-2)
    b = torch.nn.Parameter(torch.randn(64, 32, device=device) * 0.1)
    opt = Adan([{"params": [a], "lr": 5e-2}, {"params": [b], "lr": 5e-3}], eps=1e-8, weight_decay=2e-5, max_grad_norm=5.0, foreach=False)
    mirror = torch.zeros(100003, 2, device=device, dtype=torch.float16)
    opt.half_mirrors[id(a)] = mirror
    p0 = [a.detach().clone(), b.detach().clone()]
    grads = []
    for step in range(4):
        ga = torch.randn_like(a) * (10.0 if step == 1 else 0.01)     # step 1 exceeds the clip norm
        gb = torch.randn_like(b)
        grads.append([ga.clone(), gb.clone()])
        a.grad = ga.clone(); b.grad = gb.clone()
        opt.step(zero_grad=True)
        assert float(a.grad.abs().sum()) == 0.0
    ref = ref_adan_steps(p0, grads, [5e-2, 5e-3])
    assert (a.detach().double() - ref[0]).abs().max().item() < 2e-6
    assert (b.detach().double() - ref[1]).abs().max().item() < 2e-6
    assert torch.equal(mirror, a.detach().half())


def test_loss_scale_and_nonfinite_skip(device):
    torch.manual_seed(1)
    a = torch.nn.Parameter(torch.randn(5000, device=device))
    b = torch.nn.Parameter(torch.randn(5000, device=device))
    b.data.copy_(a.data)
    oa = Adan([a], lr=1e-2, max_grad_norm=5.0); ob = Adan([b], lr=1e-2, max_grad_norm=5.0)
    ob.loss_scale = 1024.0
    g = torch.randn(5000, device=device)
    a.grad = g.clone(); b.grad = g * 1024.0
    oa.step(); ob.step()
    assert (a - b).abs().max().item() < 1e-6
    before = b.detach().clone()
    b.grad = torch.full_like(b, float("inf"))
    ob.step()
    assert torch.equal(b.detach(), before)
