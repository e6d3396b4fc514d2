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


def test_nonfinite_step_clears_gradients_and_training_recovers(device):
    """a skipped step must still clear .grad when zero_grad is set — the fused backward ACCUMULATES into the buffers, so a
    stale inf would poison every later step — and must not advance the step counter: the next finite step is the optimiser's FIRST
    step (bias corrections of step 1, neg_pre_grad initialised to -g), exactly as GradScaler.step() not calling optimizer.step()."""
    torch.manual_seed(3)
    p = torch.nn.Parameter(torch.randn(4099, device=device))
    q = torch.nn.Parameter(p.detach().clone())
    skipped = Adan([p], lr=1e-2, max_grad_norm=5.0, weight_decay=2e-5)
    clean = Adan([q], lr=1e-2, max_grad_norm=5.0, weight_decay=2e-5)
    before = p.detach().clone()
    p.grad = torch.randn_like(p)
    p.grad[17] = float("inf")
    skipped.step(zero_grad=True)
    assert torch.equal(p.detach(), before), "a non-finite step must not move the parameters"
    assert float(p.grad.abs().sum()) == 0.0, "gradients must be cleared even when the step is skipped"
    for s in range(3):
        g = torch.randn(4099, device=device, generator=torch.Generator(device=device).manual_seed(s))
        p.grad.add_(g)                       # accumulate like the fused field backward does
        q.grad = g.clone()
        skipped.step(zero_grad=True)
        clean.step(zero_grad=True)
        assert torch.isfinite(p).all()
        assert (p.detach() - q.detach()).abs().max().item() < 1e-6, f"step {s}: skipped-then-recovered trajectory differs from a clean one"
    skipped.sync_steps()
    assert skipped.param_groups[0]["step"] == 3


def test_every_group_advances_and_late_gradients(device):
    """a group whose parameters have no gradient yet still counts steps like the reference (optimizer.py:191-194),
    so when its first gradient arrives at call k it is updated with the bias corrections of step k and a freshly initialised neg_pre_grad"""
    torch.manual_seed(4)
    a = torch.nn.Parameter(torch.randn(300, device=device)); b = torch.nn.Parameter(torch.randn(300, device=device))
    b0 = b.detach().clone()
    opt = Adan([{"params": [a]}, {"params": [b]}], lr=1e-2, max_grad_norm=0.0, weight_decay=0.0)
    a.grad = torch.randn_like(a)
    opt.step()
    opt.step()
    gb = torch.randn_like(b)
    b.grad = gb.clone()
    opt.step()                               # third call: b's first gradient
    b1_, b2_, b3_ = 0.98, 0.92, 0.99
    g = gb.double()
    # reference at group step 3 with fresh state: 'neg_pre_grad' not in state -> initialised to -g (optimizer.py:164), so diff = 0
    m = (1 - b1_) * g
    n = (1 - b3_) * g * g
    denom = n.sqrt() / (1 - b3_ ** 3) ** 0.5 + 1e-8
    exp = b0.double() - 1e-2 / (1 - b1_ ** 3) * m / denom
    assert (b.detach().double() - exp).abs().max().item() < 2e-6


def test_ema_matches_torch_ema_semantics(device):
    """torch_ema.ExponentialMovingAverage (nerf/utils.py:282-283,1090-1091): shadow -= (1 - d)(shadow - p), d = min(decay, (1 + n)/(10 + n));
    both the stand-alone update and the one folded into the Adan pass"""
    torch.manual_seed(5)
    p = torch.nn.Parameter(torch.randn(10007, device=device))
    opt = Adan([p], lr=1e-2)
    opt.ema_attach(0.95)
    shadow = p.detach().double().clone()
    for n in range(1, 6):
        p.grad = torch.randn_like(p)
        fused = n % 2 == 0
        opt.step(zero_grad=True, ema=fused)
        if not fused:
            opt.ema_update()
        d = min(0.95, (1 + n) / (10 + n))
        shadow = shadow - (1 - d) * (shadow - p.detach().double())
        assert (opt.ema_shadow[id(p)].double() - shadow).abs().max().item() < 1e-6, n
