"""GPU: the fused radiance-field kernels (csrc/fused_field{,_bwd}.cu) through sdf_b200.ngp.InstantNGP — self-consistency checks that need
no reference: (i) the analytic table / MLP gradients of the fused backward against central finite differences of the fused forward in
float64-accumulated losses, (ii) forward invariants (albedo mode returns no normal, normals are unit or zero, colours in [0,1], outside-box
clamping), (iii) the fp16 table mirror path equals the cast-per-call path.

The comparisons against the reference's own NeRFNetwork (its unmodified nerf/network_grid.py on its own CUDA extensions and on the
drop-in ops, points over the whole box, all four shading modes, forward and gradients) live in tests/test_gpu_dropin_reference.py."""
import numpy as np
import pytest
import torch

from sdf_b200.ngp import InstantNGP
from sdf_b200.options import default_opt

pytestmark = pytest.mark.gpu


def make_model(device, seed=0, emb_scale=0.1):
    torch.manual_seed(seed)
    m = InstantNGP(default_opt()).to(device)
    m.encoder.embeddings.data.uniform_(-emb_scale, emb_scale)
    m.invalidate_mirror()
    return m


def sample_points(device, M, seed=0, box=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    if box:
        x = torch.rand(M, 3, generator=g) * 2 - 1
    else:
        v = torch.randn(M, 3, generator=g)
        x = v / v.norm(dim=-1, keepdim=True) * (0.05 + 0.30 * torch.rand(M, 1, generator=g))
    l = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    return x.to(device), l.to(device)


@pytest.mark.parametrize("shading", ["albedo", "lambertian", "textureless", "normal"])
def test_forward_invariants(device, shading):
    m = make_model(device)
    M = 5000 + 7
    x, l = sample_points(device, M, box=True)
    x[-8:] = 1.0
    with torch.no_grad():
        s, c, n = m(x, None, l, ratio=0.3, shading=shading)
    assert s.shape == (M,) and c.shape == (M, 3) and torch.isfinite(s).all() and torch.isfinite(c).all() and (s >= 0).all()
    assert c.min().item() >= 0.0 and c.max().item() <= 1.0 + 1e-3
    if shading == "albedo":
        assert n is None
    else:
        nn_ = n.norm(dim=-1)
        assert ((nn_ - 1).abs() < 1e-3).logical_or(nn_ < 1e-6).all()
        if shading == "normal":
            assert torch.allclose(c, (n + 1) / 2, atol=1e-6)


def test_mirror_path_equals_cast_path(device):
    m = make_model(device, seed=4)
    x, l = sample_points(device, 3000, seed=4)
    with torch.no_grad():
        a = m(x, None, l, ratio=0.5, shading="lambertian")
        from sdf_b200.optimizer import Adan
        opt = Adan(m.get_params(1e-3))
        m.attach_half_mirror(opt)
        b = m(x, None, l, ratio=0.5, shading="lambertian")
    for u, v in zip(a, b):
        assert torch.equal(u, v)


@pytest.mark.parametrize("shading", ["albedo", "lambertian"])
def test_backward_against_finite_differences_of_the_forward(device, shading):
    """directional derivative of L = <g_sigma, sigma> + <g_color, color> along the gradient direction in (table, MLP) space: analytic gradient
    of the fused backward vs a central difference of the fused forward.  The forward rounds features / logits to fp16 (it mirrors the -O
    autocast graph), so the difference quotient uses a step large enough to dominate that rounding; agreement 15 %."""
    m = make_model(device, seed=2, emb_scale=0.2)
    M = 4096
    x, l = sample_points(device, M, seed=2)
    g = torch.Generator(device="cpu").manual_seed(5)
    gs = (torch.randn(M, generator=g) * 0.01).to(device)
    gc = torch.randn(M, 3, generator=g).to(device)

    def loss_of():
        s, c, _ = m(x, None, l, ratio=0.3, shading=shading)
        return (s.double() * gs.double()).sum() + (c.double() * gc.double()).sum()

    for p in m.parameters():
        p.grad = None
    loss_of().backward()
    names = ["encoder.embeddings"] + [f"sigma_net.net.{i}.weight" for i in range(3)]
    named = dict(m.named_parameters())
    for nm in names:
        p = named[nm]
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().max().item() > 0
        # along the gradient itself: the largest possible signal against the fp16 rounding noise of the forward (a random direction's
        # directional derivative is ~1/sqrt(#entries) of it and drowns in that noise)
        d = p.grad.detach().clone()
        d = d / (d.norm() + 1e-12)
        ana = (p.grad.double() * d.double()).sum().item()
        eps = 2e-2 * max(p.detach().abs().max().item(), 1e-3) / d.abs().max().item()       # largest component moves by 2 % of the parameter scale
        with torch.no_grad():
            p.add_(eps * d)
            m.invalidate_mirror()
            lp = loss_of().item()
            p.sub_(2 * eps * d)
            lm = loss_of().item()
            p.add_(eps * d)
        num = (lp - lm) / (2 * eps)
        tol = 0.15
        print(f"{shading:10s} {nm:24s} analytic {ana:+.5e} numeric {num:+.5e}")
        assert abs(ana - num) <= tol * max(abs(ana), abs(num)) + 1e-4, (nm, ana, num)


def test_density_partial_levels_zero_fill(device):
    """max_level < 1 computes only the first ceil(max_level * L) levels (gridencoder/grid.py:42,53): evaluating with max_level must equal
    evaluating a table whose remaining levels are zero"""
    m = make_model(device, seed=1)
    x, _ = sample_points(device, 4000, seed=1, box=True)
    off = m.encoder.offsets.tolist()
    for ml in (0.5, 0.26):
        L = 16
        active = max(min(int(np.ceil(ml * L)), L), 1)
        m.max_level = ml
        with torch.no_grad():
            a = m.density(x)["sigma"]
            m.max_level = None
            saved = m.encoder.embeddings.detach().clone()
            m.encoder.embeddings.data[off[active]:] = 0
            m.invalidate_mirror()
            b = m.density(x)["sigma"]
            m.encoder.embeddings.data.copy_(saved)
            m.invalidate_mirror()
        assert torch.allclose(a, b, rtol=1e-6, atol=0), (ml, (a - b).abs().max().item())
