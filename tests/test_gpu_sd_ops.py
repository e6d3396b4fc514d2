"""GPU: memory-bound SD kernels (csrc/sd_ops.cu) against fp32 PyTorch references of the same ops on the same fp16 inputs.
Tolerance: one fp16 rounding of the output (rtol 2e-3) plus fp32 statistics (atol 2e-3)."""
import math

import pytest
import torch
import torch.nn.functional as F

from sdf_b200 import _lib

pytestmark = pytest.mark.gpu


def rnd(*shape, device, scale=1.0, seed=0, shift=0.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale + shift).to(device).half()


def close(a, b, rtol=2e-3, atol=2e-3):
    a, b = a.float(), b.float()
    assert torch.isfinite(a).all()
    err = (a - b).abs() - (atol + rtol * b.abs())
    assert err.max().item() <= 0, f"max violation {err.max().item():.3g}, max abs err {(a - b).abs().max().item():.3g}"


@pytest.mark.parametrize("Nimg,HW,C,act", [(2, 4096, 320, 1), (2, 256, 1280, 0), (1, 16384, 128, 1), (2, 64, 2560, 1), (1, 1024, 960, 1),
                                           (2, 4096, 640, 1), (1, 65536, 128, 1), (2, 4096, 960, 0), (3, 100, 64, 1)])
def test_groupnorm_fwd_bwd(device, Nimg, HW, C, act):
    x = rnd(Nimg, HW, C, device=device, seed=1, shift=0.3)
    gamma = torch.randn(C, device=device) * 0.5 + 1
    beta = torch.randn(C, device=device) * 0.2
    y = torch.empty_like(x)
    stats = torch.empty(Nimg, 32, 2, device=device)
    _lib.call("sdf_groupnorm_forward", _lib.ptr(x), C, _lib.ptr(y), C, Nimg, HW, C, 32, _lib.ptr(gamma), _lib.ptr(beta), 1e-5, act, _lib.ptr(stats), _lib.stream())
    xr = x.float().permute(0, 2, 1).reshape(Nimg, C, HW).requires_grad_(True)
    yr = F.group_norm(xr, 32, gamma, beta, 1e-5)
    if act:
        yr = F.silu(yr)
    close(y, yr.permute(0, 2, 1))
    dy = rnd(Nimg, HW, C, device=device, seed=2)
    (gx,) = torch.autograd.grad(yr, xr, dy.float().permute(0, 2, 1))
    dx = torch.empty_like(x)
    bstats = torch.empty(Nimg, 32, 2, device=device)
    _lib.call("sdf_groupnorm_backward", _lib.ptr(x), C, _lib.ptr(dy), C, _lib.ptr(dx), C, Nimg, HW, C, 32, _lib.ptr(gamma), _lib.ptr(beta), 1e-5, act,
              _lib.ptr(stats), _lib.ptr(bstats), 0, _lib.stream())
    close(dx, gx.permute(0, 2, 1), rtol=4e-3, atol=4e-3)
    # accumulate mode adds onto the existing buffer
    _lib.call("sdf_groupnorm_backward", _lib.ptr(x), C, _lib.ptr(dy), C, _lib.ptr(dx), C, Nimg, HW, C, 32, _lib.ptr(gamma), _lib.ptr(beta), 1e-5, act,
              _lib.ptr(stats), _lib.ptr(bstats), 1, _lib.stream())
    close(dx, 2 * gx.permute(0, 2, 1), rtol=6e-3, atol=8e-3)


def test_layernorm_geglu_softmax(device):
    x = rnd(777, 640, device=device, seed=3, shift=-0.2)
    g = torch.randn(640, device=device) * 0.5 + 1; b = torch.randn(640, device=device) * 0.1
    y = torch.empty_like(x)
    _lib.call("sdf_layernorm_forward", _lib.ptr(x), 640, _lib.ptr(y), 640, 777, 640, _lib.ptr(g), _lib.ptr(b), 1e-5, _lib.stream())
    close(y, F.layer_norm(x.float(), (640,), g, b, 1e-5))
    z = rnd(300, 2 * 1280, device=device, seed=4)
    o = torch.empty(300, 1280, device=device, dtype=torch.float16)
    _lib.call("sdf_geglu", _lib.ptr(z), 2560, _lib.ptr(o), 1280, 300, 1280, _lib.stream())
    close(o, z[:, :1280].float() * F.gelu(z[:, 1280:].float()))
    for rows, cols, ld in [(64, 4096, 4096), (500, 77, 128), (33, 1024, 1024)]:
        s = rnd(rows, ld, device=device, seed=5, scale=3.0)
        p = torch.zeros_like(s)
        _lib.call("sdf_softmax_rows", _lib.ptr(s), _lib.ptr(p), rows, cols, ld, 0.7, _lib.stream())
        pr = torch.softmax(s[:, :cols].float() * 0.7, -1)
        close(p[:, :cols], pr, atol=1e-4)
        assert (p[:, cols:] == 0).all()
        if cols % 8 == 0:
            dp = rnd(rows, ld, device=device, seed=6)
            ds = torch.empty_like(s)
            _lib.call("sdf_softmax_rows_backward", _lib.ptr(p), _lib.ptr(dp), _lib.ptr(ds), rows, cols, ld, 0.7, _lib.stream())
            pf = p.float(); dpf = dp.float()
            ref = 0.7 * pf * (dpf - (pf * dpf).sum(-1, keepdim=True))
            close(ds, ref, atol=1e-4)


def test_resample_im2col_glue(device):
    x = rnd(2, 8, 8, 64, device=device, seed=7)
    y = torch.empty(2, 16, 16, 64, device=device, dtype=torch.float16)
    _lib.call("sdf_upsample_nearest2", _lib.ptr(x), 64, _lib.ptr(y), 64, 2, 8, 8, 64, _lib.stream())
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(y.float(), ref)
    # stride-2 conv via im2col: UNet (pad 1 all sides) and VAE ((0,1,0,1) pad) conventions, then adjoint identity <col, im2col(x)> = <col2im(col), x>
    for (pt, pl, Ho) in [(1, 1, 8), (0, 0, 8)]:
        a = rnd(2, 16, 16, 64, device=device, seed=8)
        col = torch.empty(2, Ho, Ho, 9 * 64, device=device, dtype=torch.float16)
        _lib.call("sdf_im2col_s2", _lib.ptr(a), 64, _lib.ptr(col), 2, 16, 16, 64, Ho, Ho, pt, pl, _lib.stream())
        xp = a.float().permute(0, 3, 1, 2)
        xp = F.pad(xp, (1, 1, 1, 1)) if pt else F.pad(xp, (0, 1, 0, 1))
        ref = F.unfold(xp, 3, stride=2).view(2, 64, 9, Ho, Ho).permute(0, 3, 4, 2, 1).reshape(2, Ho, Ho, 9 * 64)
        assert torch.equal(col.float(), ref)
        dcol = rnd(2, Ho, Ho, 9 * 64, device=device, seed=9)
        dx = torch.empty_like(a)
        _lib.call("sdf_col2im_s2", _lib.ptr(dcol), _lib.ptr(dx), 64, 2, 16, 16, 64, Ho, Ho, pt, pl, _lib.stream())
        lhs = (dcol.double() * col.double()).sum().item()
        rhs = (dx.double() * a.double()).sum().item()
        assert abs(lhs - rhs) < 2e-2 * max(1.0, abs(lhs)), (lhs, rhs)
    t = torch.tensor([20, 500, 980], dtype=torch.int32, device=device)
    e = torch.empty(3, 320, device=device, dtype=torch.float16)
    _lib.call("sdf_timestep_embedding", _lib.ptr(t), 3, 320, _lib.ptr(e), 320, _lib.stream())
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=device) / half)
    args = t[:, None].float() * freqs[None]
    close(e, torch.cat([torch.cos(args), torch.sin(args)], -1), rtol=1e-3, atol=2e-3)
    a = rnd(100, 96, device=device, seed=10); b = rnd(100, 64, device=device, seed=11); o = torch.zeros(100, 160, device=device, dtype=torch.float16)
    _lib.call("sdf_copy2d", _lib.ptr(a), 96, _lib.ptr(o), 160, 100, 96, _lib.stream())
    _lib.call("sdf_copy2d", _lib.ptr(b), 64, o.data_ptr() + 96 * 2, 160, 100, 64, _lib.stream())
    assert torch.equal(o, torch.cat([a, b], -1))
    s = torch.empty(100, 64, device=device, dtype=torch.float16)
    _lib.call("sdf_add2d", _lib.ptr(a), 96, _lib.ptr(b), 64, _lib.ptr(s), 64, 100, 64, _lib.stream())
    close(s, a[:, :64].float() + b.float())
    tr = torch.empty(3, 96, 104, device=device, dtype=torch.float16)
    a3 = rnd(3, 100, 96, device=device, seed=12)
    _lib.call("sdf_transpose2d", _lib.ptr(a3), 96, _lib.ptr(tr), 104, 3, 100, 96, _lib.stream())
    assert torch.equal(tr[:, :, :100], a3.permute(0, 2, 1))


@pytest.mark.parametrize("B,heads,n,nkv,d,fused", [(2, 8, 4096, 4096, 40, True), (2, 8, 1024, 1024, 80, True), (2, 8, 256, 256, 160, True),
                                                    (2, 8, 64, 64, 160, True), (2, 8, 4096, 77, 40, False), (2, 8, 256, 77, 160, False),
                                                    (1, 4, 100, 5, 32, False), (1, 2, 70, 130, 64, False)])
def test_flash_attention(device, B, heads, n, nkv, d, fused):
    """csrc/flash_attn.cu vs softmax(q k^T / sqrt(d)) v in fp32 on the same fp16 inputs (ldm/modules/attention.py:170-193).
    P is rounded to fp16 before the P V product (as the reference does under autocast): atol 2e-3 of max|v|."""
    C = heads * d
    if fused:                                  # q|k|v side by side as the fused projection GEMM writes them
        qkv = rnd(B, n, 3 * C, device=device, seed=3, scale=1.5)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:
        q = rnd(B, n, C, device=device, seed=4, scale=1.5)
        kv = rnd(B, nkv, 2 * C, device=device, seed=5, scale=1.5)
        k, v = kv[..., :C], kv[..., C:]
    o = torch.zeros(B, n, C, device=device, dtype=torch.float16)
    _lib.call("sdf_flash_attention", q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, heads, n, nkv, d,
              q.stride(1), k.stride(1), C, d ** -0.5, _lib.stream())
    qf, kf, vf = (t.float().reshape(B, -1, heads, d).permute(0, 2, 1, 3) for t in (q, k, v))
    ref = torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, -1) @ vf
    ref = ref.permute(0, 2, 1, 3).reshape(B, n, C)
    close(o, ref, rtol=4e-3, atol=4e-3)


@pytest.mark.parametrize("Nimg,H,W,Cin,C", [(1, 64, 64, 3, 128), (2, 17, 23, 3, 128), (1, 9, 8, 4, 256), (1, 1, 1, 3, 128)])
def test_conv3x3_small_cin_and_its_data_gradient(device, Nimg, H, W, Cin, C):
    """the VAE's conv_in (ldm/modules/diffusionmodules/model.py:387) as a direct convolution, against F.conv2d / its autograd on the same fp16 image"""
    x = torch.zeros(Nimg, H, W, 8, device=device, dtype=torch.float16)
    x[..., :Cin] = rnd(Nimg, H, W, Cin, device=device, seed=31)
    g = torch.Generator(device="cpu").manual_seed(32)
    w = (torch.randn(C, Cin, 3, 3, generator=g) * 0.2).to(device)
    b = (torch.randn(C, generator=g) * 0.1).to(device)
    y = torch.empty(Nimg, H, W, C, device=device, dtype=torch.float16)
    _lib.call("sdf_conv3x3_small_cin_forward", _lib.ptr(x), 8, _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), C, Nimg, H, W, Cin, C, _lib.stream())
    xr = x[..., :Cin].float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.conv2d(xr, w, b, padding=1)
    close(y, yr.permute(0, 2, 3, 1))
    dy = rnd(Nimg, H, W, C, device=device, seed=33, scale=0.1)
    dx = torch.full((Nimg, H, W, 8), 7.0, device=device, dtype=torch.float16)
    _lib.call("sdf_conv3x3_small_cin_dgrad", _lib.ptr(dy), C, _lib.ptr(w), _lib.ptr(dx), 8, Nimg, H, W, Cin, C, _lib.stream())
    yr.backward(dy.float().permute(0, 3, 1, 2))
    close(dx[..., :Cin], xr.grad.permute(0, 2, 3, 1), atol=4e-3)
    assert (dx[..., Cin:4] == 0).all() and (dx[..., 4:] == 7.0).all()        # writes exactly one 4-channel pixel
    with pytest.raises(RuntimeError):
        _lib.call("sdf_conv3x3_small_cin_forward", _lib.ptr(x), 8, _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), C, Nimg, H, W, 5, C, _lib.stream())
