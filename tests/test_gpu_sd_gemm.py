"""GPU: the tcgen05 implicit-GEMM kernel (csrc/sd_gemm.cu) against an fp32 PyTorch reference of the same op
(F.conv2d / F.linear on the same fp16-rounded operands).  Tolerance: fp16 inputs, fp32 accumulate, fp16 output ->
rtol 2e-3 + atol 2e-3 * sqrt(K/64) on O(1) data."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from sdf_b200 import _lib

pytestmark = pytest.mark.gpu


def run_plan(a_nhwc, w_oihw, bias=None, temb=None, residual=None, act=0, alpha=1.0, splitk=1, block_n=128, n_valid=None, lda_extra=0):
    """a_nhwc [Nimg,H,W,Cin] fp16 cuda; w [Cout,Cin,kh,kw] fp16 -> out [Nimg,H,W,Cout] fp16 via the C ABI"""
    dev = a_nhwc.device
    Nimg, H, W, Cin = a_nhwc.shape
    Cout, _, kh, kw = w_oihw.shape
    taps = kh * kw
    if lda_extra:
        buf = torch.zeros(Nimg, H, W, Cin + lda_extra, device=dev, dtype=torch.float16)
        buf[..., :Cin] = a_nhwc
        a_used, lda = buf, Cin + lda_extra
    else:
        a_used, lda = a_nhwc.contiguous(), Cin
    rows = ((Cout + block_n - 1) // block_n) * block_n
    wt = torch.zeros(rows, taps * Cin, device=dev, dtype=torch.float16)
    wt[:Cout] = w_oihw.permute(0, 2, 3, 1).reshape(Cout, taps * Cin)
    N = Cout if n_valid is None else n_valid
    ldo = ((N + 7) // 8) * 8
    out = torch.full((Nimg * H * W, ldo), float("nan"), device=dev, dtype=torch.float16)
    ws = torch.empty(Nimg * H * W, N, device=dev, dtype=torch.float32) if splitk > 1 else None
    res_t = None
    if residual is not None:
        res_t = residual.reshape(Nimg * H * W, -1).contiguous()
    plan = _lib.lib().cdll.sdf_gemm_plan_create(
        _lib.ptr(a_used), lda, _lib.ptr(wt), rows, Nimg, H, W, Cin, taps, N, _lib.ptr(out), ldo, _lib.ptr(bias),
        _lib.ptr(temb), 0 if temb is None else temb.shape[1], _lib.ptr(res_t), 0 if res_t is None else res_t.shape[1],
        act, alpha, splitk, _lib.ptr(ws), block_n)
    assert plan >= 0, _lib.lib().last_error()
    _lib.call("sdf_gemm_run", plan, _lib.stream())
    torch.cuda.synchronize()
    _lib.call("sdf_gemm_plan_destroy", plan)
    return out[:, :N].reshape(Nimg, H, W, N)


def ref_conv(a_nhwc, w, bias=None, temb=None, residual=None, act=0, alpha=1.0):
    x = a_nhwc.float().permute(0, 3, 1, 2)
    y = F.conv2d(x, w.float(), None, padding=w.shape[-1] // 2) * alpha
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    if temb is not None:
        y = y + temb.float()[:, :, None, None]
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.float()
    if act == 1:
        y = F.silu(y)
    elif act == 2:
        y = F.gelu(y)
    return y


def check(out, ref, K):
    out = out.float()
    assert torch.isfinite(out).all(), "non-finite / unwritten outputs"
    tol = 2e-3 * ref.abs() + 2e-3 * math.sqrt(K / 64) * max(1.0, ref.abs().max().item() / 4)
    bad = ((out - ref).abs() > tol)
    assert not bad.any(), f"{int(bad.sum())} / {bad.numel()} mismatches, max err {(out - ref).abs().max().item():.4g} (ref max {ref.abs().max().item():.3g})"


def rnd(*shape, device, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(device).half()


@pytest.mark.parametrize("M,K,N,bn", [(8192, 320, 320, 160), (154, 768, 320, 160), (2, 320, 1280, 160), (512, 1280, 1280, 128),
                                      (4096, 64, 64, 64), (300, 128, 72, 64)])
def test_linear(device, M, K, N, bn):
    a = rnd(1, 1, M, K, device=device, seed=1)
    w = rnd(N, K, 1, 1, device=device, scale=1 / math.sqrt(K), seed=2)
    bias = rnd(N, device=device, seed=3).float()
    out = run_plan(a, w, bias=bias, block_n=bn)
    check(out, ref_conv(a, w, bias=bias), K)


@pytest.mark.parametrize("Nimg,H,W,Cin,Cout,bn", [(2, 64, 64, 320, 320, 160), (2, 32, 32, 640, 640, 160), (2, 16, 16, 128, 256, 128),
                                                   (2, 8, 8, 256, 320, 160), (1, 128, 256, 64, 128, 128), (3, 8, 8, 64, 64, 64)])
def test_conv3x3(device, Nimg, H, W, Cin, Cout, bn):
    a = rnd(Nimg, H, W, Cin, device=device, seed=4)
    w = rnd(Cout, Cin, 3, 3, device=device, scale=1 / math.sqrt(9 * Cin), seed=5)
    bias = rnd(Cout, device=device, seed=6).float()
    out = run_plan(a, w, bias=bias, block_n=bn)
    check(out, ref_conv(a, w, bias=bias), 9 * Cin)


def test_epilogue_variants_and_strided_input(device):
    Nimg, H, W, Cin, Cout = 2, 32, 32, 128, 320
    a = rnd(Nimg, H, W, Cin, device=device, seed=7)
    w = rnd(Cout, Cin, 3, 3, device=device, scale=1 / math.sqrt(9 * Cin), seed=8)
    bias = rnd(Cout, device=device, seed=9).float()
    temb = rnd(Nimg, Cout, device=device, seed=10)
    res = rnd(Nimg, H, W, Cout, device=device, seed=11)
    for act in (0, 1, 2):
        out = run_plan(a, w, bias=bias, temb=temb, residual=res, act=act, alpha=0.5, block_n=160, lda_extra=64)
        check(out, ref_conv(a, w, bias=bias, temb=temb, residual=res, act=act, alpha=0.5), 9 * Cin)
    # 1x1 conv, output-channel tail (N = 4 valid of a padded weight tile), no bias
    w1 = rnd(4, Cin, 1, 1, device=device, scale=1 / math.sqrt(Cin), seed=12)
    out = run_plan(a, w1, block_n=64)
    check(out, ref_conv(a, w1), Cin)


@pytest.mark.parametrize("splitk", [2, 5, 16])
def test_splitk(device, splitk):
    Nimg, H, W, Cin, Cout = 2, 8, 8, 1280, 1280
    a = rnd(Nimg, H, W, Cin, device=device, seed=13)
    w = rnd(Cout, Cin, 3, 3, device=device, scale=1 / math.sqrt(9 * Cin), seed=14)
    bias = rnd(Cout, device=device, seed=15).float()
    res = rnd(Nimg, H, W, Cout, device=device, seed=16)
    out = run_plan(a, w, bias=bias, residual=res, act=1, splitk=splitk, block_n=160)
    check(out, ref_conv(a, w, bias=bias, residual=res, act=1), 9 * Cin)
