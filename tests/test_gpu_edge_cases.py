"""GPU: degenerate sizes through the drop-in packages / C ABI — empty ray sets, empty sample sets, single-element inputs, ragged tails.
The reference's kernels are launched with zero-sized grids in these cases (and fail); the contract here is 'returns empty / zero
outputs, never touches memory, never raises'."""
import numpy as np
import pytest
import torch

from sdf_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def test_empty_rays_and_samples(device):
    import raymarching
    z3 = torch.zeros(0, 3, device=device)
    aabb = torch.tensor([-1., -1, -1, 1, 1, 1], device=device)
    nears, fars = raymarching.near_far_from_aabb(z3, z3, aabb, 0.2)
    assert nears.shape == (0,) and fars.shape == (0,)
    bf = torch.from_numpy(synth.occupancy_bitfield("blob", 128, 1, 1.0, seed=0)).to(device)
    xyzs, dirs, ts, rays = raymarching.march_rays_train(z3, z3, 1.0, bf, 1, 128, nears, fars)
    assert xyzs.shape == (0, 3) and ts.shape == (0, 2) and rays.shape == (0, 2)
    w, ws, dep, img = raymarching.composite_rays_train(torch.zeros(0, device=device, requires_grad=True), torch.zeros(0, 3, device=device), ts, rays)
    assert w.shape == (0,) and img.shape == (0, 3)
    assert raymarching.flatten_rays(rays, 0).shape == (0,)
    assert raymarching.morton3D(torch.zeros(0, 3, dtype=torch.int32, device=device)).shape == (0,)


def test_empty_encoders_and_field(device):
    import freqencoder
    import gridencoder
    import shencoder
    from sdf_b200.ngp import InstantNGP
    from sdf_b200.options import default_opt
    enc = gridencoder.GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048,
                                  gridtype='hash', align_corners=False, interpolation='smoothstep').to(device)
    x0 = torch.zeros(0, 3, device=device)
    with torch.autocast("cuda", dtype=torch.float16):
        assert enc(x0, bound=1).shape == (0, 32)
    assert freqencoder.FreqEncoder(input_dim=3, degree=6).to(device)(x0).shape == (0, 39)
    assert shencoder.SHEncoder(input_dim=3, degree=4).to(device)(x0).shape == (0, 16)
    net = InstantNGP(default_opt(h=8, w=8)).to(device)
    s, c, n = net(x0, None, torch.zeros(0, 3, device=device), ratio=0.5, shading="lambertian")
    assert s.shape == (0,) and c.shape == (0, 3)
    (s.sum() + c.sum()).backward()                      # empty backward: gradients stay zero / None, no launch error
    torch.cuda.synchronize()
    # one single sample, every shading mode
    x1 = torch.tensor([[0.1, -0.2, 0.05]], device=device)
    for sh in ("albedo", "lambertian", "textureless", "normal"):
        s, c, n = net(x1, None, torch.tensor([[0., 0., 1.]], device=device), ratio=0.3, shading=sh)
        assert torch.isfinite(s).all() and torch.isfinite(c).all()
        (s.sum() + c.sum()).backward()
    torch.cuda.synchronize()


@pytest.mark.parametrize("n,nkv,d", [(1, 1, 40), (5, 3, 80), (65, 129, 40), (130, 64, 160)])
def test_flash_attention_ragged(device, n, nkv, d):
    """query / key counts that are not multiples of the 64-wide tiles, including a single key (softmax of one element = 1)"""
    B, heads = 2, 2
    C = heads * d
    g = torch.Generator(device="cpu").manual_seed(n * 1000 + nkv)
    q = torch.randn(B, n, C, generator=g).to(device).half()
    kv = torch.randn(B, nkv, 2 * C, generator=g).to(device).half()
    k, v = kv[..., :C], kv[..., C:]
    o = torch.full((B, n, C), float("nan"), device=device, dtype=torch.float16)
    _lib.call("sdf_flash_attention", q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, heads, n, nkv, d, C, 2 * C, C, d ** -0.5, _lib.stream())
    qf, kf, vf = (t.float().reshape(B, -1, heads, d).permute(0, 2, 1, 3) for t in (q, k, v))
    ref = (torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, -1) @ vf).permute(0, 2, 1, 3).reshape(B, n, C)
    assert torch.isfinite(o.float()).all()
    assert (o.float() - ref).abs().max().item() < 6e-3 * max(1.0, ref.abs().max().item())


def test_adan_without_gradients_and_zero_size(device):
    from sdf_b200.optimizer import Adan
    p = [torch.nn.Parameter(torch.randn(10, device=device)), torch.nn.Parameter(torch.zeros(0, device=device))]
    opt = Adan([{"params": p, "lr": 1e-2}], eps=1e-8, weight_decay=2e-5, max_grad_norm=5.0, foreach=False)
    before = p[0].detach().clone()
    opt.step()                                          # no .grad anywhere: a no-op, not an error
    assert torch.equal(p[0].detach(), before)
    p[0].grad = torch.ones_like(p[0]); p[1].grad = torch.zeros_like(p[1])
    opt.step()
    assert torch.isfinite(p[0]).all() and not torch.equal(p[0].detach(), before)


@pytest.mark.parametrize("n,d,heads", [(4096, 40, 8), (1024, 40, 2), (600, 40, 3), (1000, 64, 2), (512, 32, 2), (777, 48, 1)])
def test_flash_attention_tcgen05_path(device, n, d, heads):
    """long self-attention takes the tcgen05 / TMEM / TMA kernel (csrc/flash_attn_tc.cu: n, nkv >= 512, d <= 64), including query / key
    counts that are not multiples of its 128-wide tiles, read in place from a fused q|k|v projection buffer (row stride 3 * heads * d);
    same bar as the mma.sync kernel: |err| <= 4e-3 + 4e-3 |ref| against fp32 softmax(q k^T / sqrt(d)) v on the same fp16 inputs"""
    B = 2
    C = heads * d
    g = torch.Generator(device="cpu").manual_seed(n + d)
    qkv = (torch.randn(B, n, 3 * C, generator=g) * 1.5).to(device).half()
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    o = torch.full((B, n, C), float("nan"), device=device, dtype=torch.float16)
    _lib.call("sdf_flash_attention", q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, heads, n, n, d, 3 * C, 3 * C, C, d ** -0.5, _lib.stream())
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().reshape(B, n, heads, d).permute(0, 2, 1, 3) for t in (q, k, v))
    ref = (torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, -1) @ vf).permute(0, 2, 1, 3).reshape(B, n, C)
    assert torch.isfinite(o.float()).all()
    err = (o.float() - ref).abs()
    assert (err <= 4e-3 + 4e-3 * ref.abs()).all(), (err.max().item(), ref.abs().max().item())


@pytest.mark.parametrize("n,d,heads", [(2048, 40, 2), (1111, 64, 1)])
def test_flash_attention_tcgen05_growing_scores(device, n, d, heads):
    """keys whose magnitude grows 12x along the sequence: the running row maximum keeps moving by many powers of two from key tile to
    key tile, which is the path that rescales the output accumulator (version 2 of the tcgen05 kernel keeps O in TMEM and only rescales it
    when a row's maximum grows by more than 2^8; version 1 rescales every tile).  Same bar as the other attention tests."""
    B = 2
    C = heads * d
    g = torch.Generator(device="cpu").manual_seed(7 * n + d)
    qkv = torch.randn(B, n, 3 * C, generator=g) * 1.5
    qkv[..., C:2 * C] *= torch.linspace(1.0, 12.0, n).view(1, n, 1)
    qkv = qkv.to(device).half()
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    o = torch.full((B, n, C), float("nan"), device=device, dtype=torch.float16)
    _lib.call("sdf_flash_attention", q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, heads, n, n, d, 3 * C, 3 * C, C, d ** -0.5, _lib.stream())
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().reshape(B, n, heads, d).permute(0, 2, 1, 3) for t in (q, k, v))
    ref = (torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, -1) @ vf).permute(0, 2, 1, 3).reshape(B, n, C)
    assert torch.isfinite(o.float()).all()
    err = (o.float() - ref).abs()
    assert (err <= 4e-3 + 4e-3 * ref.abs()).all(), (err.max().item(), ref.abs().max().item())
