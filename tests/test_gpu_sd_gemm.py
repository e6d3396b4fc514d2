"""GPU: the tcgen05 implicit-GEMM kernel (csrc/sd_gemm.cu) against an fp32 PyTorch reference of the same op
(F.conv2d / F.linear / einsum on the same fp16-rounded operands).  Tolerance: fp16 inputs, fp32 accumulate, fp16 output ->
rtol 2e-3 + atol 2e-3 * sqrt(K/64) on O(1) data."""
import math

import pytest
import torch
import torch.nn.functional as F

from sdf_b200 import gemm

pytestmark = pytest.mark.gpu


def run_conv(a_nhwc, w_oihw, bias=None, temb=None, residual=None, act=None, alpha=1.0, splitk=1, block_n=128, lda_extra=0, cta_pair=0):
    dev = a_nhwc.device
    Nimg, H, W, Cin = a_nhwc.shape
    Cout, _, kh, kw = w_oihw.shape
    if lda_extra:
        buf = torch.full((Nimg, H, W, Cin + lda_extra), 7.0, device=dev, dtype=torch.float16)   # junk beyond c_valid must be ignored
        buf[..., :Cin] = a_nhwc
        a_used = buf
    else:
        a_used = a_nhwc.contiguous()
    wt = gemm.pack_conv_weight(w_oihw)
    ldo = ((Cout + 7) // 8) * 8
    out = torch.full((Nimg, H, W, ldo), float("nan"), device=dev, dtype=torch.float16)
    plan = gemm.conv_plan(a_used, Cin, wt, Cout, out, taps=kh * kw, bias=bias, temb=temb, residual=residual, act=act, alpha=alpha,
                          splitk=splitk, block_n=block_n, cta_pair=cta_pair)
    plan.run()
    torch.cuda.synchronize()
    return out[..., :Cout]


def ref_conv(a_nhwc, w, bias=None, temb=None, residual=None, act=None, alpha=1.0):
    x = a_nhwc.float().permute(0, 3, 1, 2)
    y = F.conv2d(x, w.float(), None, padding=w.shape[-1] // 2) * alpha
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    if temb is not None:
        y = y + temb.float()[:, :, None, None]
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.float()
    if act == "silu":
        y = F.silu(y)
    elif act == "gelu":
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
                                      (4096, 64, 64, 64), (300, 128, 72, 64), (77, 40, 24, 64)])
def test_linear(device, M, K, N, bn):
    a = rnd(1, 1, M, K, device=device, seed=1)
    w = rnd(N, K, 1, 1, device=device, scale=1 / math.sqrt(K), seed=2)
    bias = rnd(N, device=device, seed=3).float()
    out = run_conv(a, w, bias=bias, block_n=bn)
    check(out, ref_conv(a, w, bias=bias), K)


@pytest.mark.parametrize("Nimg,H,W,Cin,Cout,bn", [(2, 64, 64, 320, 320, 160), (2, 32, 32, 640, 640, 160), (2, 16, 16, 128, 256, 128),
                                                   (2, 8, 8, 256, 320, 160), (1, 128, 256, 64, 128, 128), (3, 8, 8, 64, 64, 64),
                                                   (1, 8, 8, 64, 64, 64), (2, 64, 64, 4, 320, 160), (1, 32, 32, 320, 4, 64)])
def test_conv3x3(device, Nimg, H, W, Cin, Cout, bn):
    a = rnd(Nimg, H, W, max(Cin, 8), device=device, seed=4)
    a[..., Cin:] = 3.0            # channels beyond c_valid exist in memory (alignment padding) but must read as zero
    w = rnd(Cout, Cin, 3, 3, device=device, scale=1 / math.sqrt(9 * Cin), seed=5)
    bias = rnd(Cout, device=device, seed=6).float()
    wt = gemm.pack_conv_weight(w)
    ldo = ((Cout + 7) // 8) * 8
    out = torch.full((Nimg, H, W, ldo), float("nan"), device=device, dtype=torch.float16)
    gemm.conv_plan(a, Cin, wt, Cout, out, taps=9, bias=bias, block_n=bn).run()
    torch.cuda.synchronize()
    check(out[..., :Cout], ref_conv(a[..., :Cin], w, bias=bias), 9 * Cin)


def test_epilogue_variants_and_strided_input(device):
    Nimg, H, W, Cin, Cout = 2, 32, 32, 128, 320
    a = rnd(Nimg, H, W, Cin, device=device, seed=7)
    w = rnd(Cout, Cin, 3, 3, device=device, scale=1 / math.sqrt(9 * Cin), seed=8)
    bias = rnd(Cout, device=device, seed=9).float()
    temb = rnd(Nimg, Cout, device=device, seed=10)
    res = rnd(Nimg, H, W, Cout, device=device, seed=11)
    for act in (None, "silu", "gelu"):
        out = run_conv(a, w, bias=bias, temb=temb, residual=res, act=act, alpha=0.5, block_n=160, lda_extra=64)
        check(out, ref_conv(a, w, bias=bias, temb=temb, residual=res, act=act, alpha=0.5), 9 * Cin)
    w1 = rnd(4, Cin, 1, 1, device=device, scale=1 / math.sqrt(Cin), seed=12)
    out = run_conv(a, w1, block_n=64)
    check(out, ref_conv(a, w1), Cin)


@pytest.mark.parametrize("splitk", [2, 5, 16, 64])
def test_splitk(device, splitk):
    Nimg, H, W, Cin, Cout = 2, 8, 8, 1280, 1280
    a = rnd(Nimg, H, W, Cin, device=device, seed=13)
    w = rnd(Cout, Cin, 3, 3, device=device, scale=1 / math.sqrt(9 * Cin), seed=14)
    bias = rnd(Cout, device=device, seed=15).float()
    res = rnd(Nimg, H, W, Cout, device=device, seed=16)
    out = run_conv(a, w, bias=bias, residual=res, act="silu", splitk=splitk, block_n=160)
    check(out, ref_conv(a, w, bias=bias, residual=res, act="silu"), 9 * Cin)


@pytest.mark.parametrize("B,heads,n,nkv,d", [(2, 8, 1024, 1024, 80), (2, 8, 256, 77, 160), (2, 8, 4096, 77, 40), (2, 8, 64, 64, 160), (1, 1, 4096, 4096, 512)])
def test_batched_attention_products(device, B, heads, n, nkv, d):
    """S = scale * Q K^T per (batch, head) straight from [tokens, heads*d] projections (K tail by TMA zero fill), then
    O = P V with V^T as the weight operand, written back head-interleaved."""
    C = heads * d
    q = rnd(B, n, C, device=device, seed=20)
    k = rnd(B, nkv, C, device=device, seed=21)
    v = rnd(B, nkv, C, device=device, seed=22)
    scale = d ** -0.5
    ld_s = ((nkv + 63) // 64) * 64
    S = torch.zeros(B, heads, n, ld_s, device=device, dtype=torch.float16)
    Kit = ((d + 63) // 64) * 64
    p1 = gemm.GemmPlan(q, (C, d, n * C), d, k, (C, d, nkv * C), d, nkv, B, heads, n, Kit, 1, nkv, S, (ld_s, n * ld_s, heads * n * ld_s),
                       alpha=scale, block_n=64 if nkv <= 64 else 128)
    p1.run()
    torch.cuda.synchronize()
    qh = q.float().view(B, n, heads, d).permute(0, 2, 1, 3)
    kh = k.float().view(B, nkv, heads, d).permute(0, 2, 1, 3)
    vh = v.float().view(B, nkv, heads, d).permute(0, 2, 1, 3)
    S_ref = torch.einsum("bhid,bhjd->bhij", qh, kh) * scale
    check(S[..., :nkv], S_ref, d)
    assert (S[..., nkv:] == 0).all()
    P = torch.softmax(S_ref, dim=-1).half()
    Pbuf = torch.zeros(B, heads, n, ld_s, device=device, dtype=torch.float16)
    Pbuf[..., :nkv] = P
    vt = torch.zeros(B, C, ld_s, device=device, dtype=torch.float16)          # V^T [b][h*d + j][kv]
    vt[..., :nkv] = v.permute(0, 2, 1)
    O = torch.full((B, n, C), float("nan"), device=device, dtype=torch.float16)
    p2 = gemm.GemmPlan(Pbuf, (ld_s, n * ld_s, heads * n * ld_s), nkv, vt, (ld_s, d * ld_s, C * ld_s), nkv, d, B, heads, n, ld_s, 1, d,
                       O, (C, d, n * C), block_n=64 if d <= 64 else (160 if d % 160 == 0 else 128))
    p2.run()
    torch.cuda.synchronize()
    O_ref = torch.einsum("bhij,bhjd->bhid", P.float(), vh).permute(0, 2, 1, 3).reshape(B, n, C)
    check(O, O_ref, nkv)


# ---- CTA-pair variant (tcgen05 cta_group::2): same contract, M = 256 per MMA, weight tile split across the two CTAs
@pytest.mark.parametrize("M,K,N,bn", [(8192, 320, 320, 160), (8192, 320, 2560, 256), (154, 768, 320, 160), (512, 1280, 1280, 256),
                                      (4096, 64, 128, 128), (300, 128, 72, 128), (129, 64, 384, 256), (100000, 192, 256, 256)])
def test_linear_cta_pair(device, M, K, N, bn):
    a = rnd(1, 1, M, K, device=device, seed=1)
    w = rnd(N, K, 1, 1, device=device, scale=1 / math.sqrt(K), seed=2)
    bias = rnd(N, device=device, seed=3).float()
    out = run_conv(a, w, bias=bias, block_n=bn, cta_pair=1)
    check(out, ref_conv(a, w, bias=bias), K)


@pytest.mark.parametrize("Nimg,H,W,Cin,Cout,bn,sk", [(2, 64, 64, 320, 320, 160, 1), (2, 32, 32, 640, 640, 160, 1), (2, 16, 16, 128, 256, 256, 1),
                                                      (2, 16, 16, 1280, 1280, 256, 4), (1, 128, 256, 64, 128, 128, 1), (3, 8, 8, 64, 128, 128, 1),
                                                      (1, 96, 96, 128, 256, 256, 1), (2, 64, 64, 4, 320, 160, 1), (3, 24, 40, 72, 200, 256, 2)])
def test_conv3x3_cta_pair(device, Nimg, H, W, Cin, Cout, bn, sk):
    a = rnd(Nimg, H, W, max(Cin, 8), device=device, seed=4)
    a[..., Cin:] = 3.0
    w = rnd(Cout, Cin, 3, 3, device=device, scale=1 / math.sqrt(9 * Cin), seed=5)
    bias = rnd(Cout, device=device, seed=6).float()
    temb = rnd(Nimg, Cout, device=device, seed=7)
    res = rnd(Nimg, H, W, Cout, device=device, seed=8)
    dev = a.device
    wt = gemm.pack_conv_weight(w)
    out = torch.full((Nimg, H, W, Cout), float("nan"), device=dev, dtype=torch.float16)
    plan = gemm.conv_plan(a, Cin, wt, Cout, out, taps=9, bias=bias, temb=temb, residual=res, act="silu", splitk=sk, block_n=bn, cta_pair=1)
    for _ in range(3):            # persistent barriers / TMEM hand-over must survive repeated launches
        plan.run()
    torch.cuda.synchronize()
    check(out, ref_conv(a[..., :Cin], w, bias=bias, temb=temb, residual=res, act="silu"), 9 * Cin)


@pytest.mark.parametrize("M,K,inner,bn,pair", [(300, 128, 160, 160, 0), (8192, 320, 1280, 256, 1), (512, 1280, 5120, 128, 0)])
def test_geglu_epilogue(device, M, K, inner, bn, pair):
    """GEGLU fused into the projection (ldm/modules/attention.py:37-45): rows interleaved [16 value | 16 gate] per 32-row chunk"""
    from sdf_b200.sd_engine import geglu_row_permutation
    a = rnd(1, 1, M, K, device=device, seed=11)
    w = rnd(2 * inner, K, device=device, scale=1 / math.sqrt(K), seed=12)
    bias = rnd(2 * inner, device=device, seed=13).float()
    perm = geglu_row_permutation(inner, device)
    wt = gemm.pack_conv_weight(w[perm].reshape(2 * inner, K, 1, 1))
    out = torch.full((1, 1, M, inner), float("nan"), device=device, dtype=torch.float16)
    plan = gemm.GemmPlan(a, (K, M * K, M * K), K, wt, (wt.shape[1], 0, 0), wt.shape[1], 2 * inner, 1, 1, M, wt.shape[1], 1, 2 * inner,
                         out, (inner, M * inner, M * inner), bias=bias[perm].contiguous(), act="geglu", block_n=bn, cta_pair=pair)
    plan.run()
    torch.cuda.synchronize()
    y = a.float().view(M, K) @ w.float().t() + bias
    ref = y[:, :inner] * F.gelu(y[:, inner:])
    check(out.view(M, inner), ref, K)


@pytest.mark.parametrize("Cout,H,block_n,pair,coff,Ctot", [(320, 64, 160, 0, 0, 320), (320, 32, 128, 0, 640, 960), (128, 64, 128, 0, 0, 128),
                                                          (256, 64, 256, 1, 0, 256), (640, 8, 128, 0, 0, 640)])
def test_epilogue_groupnorm_statistics(device, Cout, H, block_n, pair, coff, Ctot):
    """sdf_gemm_plan_set_gn_stats: per-(image, group) sum / sum of squares of the fp16 outputs, accumulated by the GEMM epilogue for a
    consumer GroupNorm(32, Ctot) whose input holds this product at channel offset `coff` (concatenated skip inputs), against torch sums of
    the stored tensor; and the apply-only GroupNorm on those statistics against F.group_norm."""
    from sdf_b200 import _lib
    Nimg, Cin = 2, 64
    g = torch.Generator(device="cpu").manual_seed(Cout + H)
    a = (torch.randn(Nimg, H, H, Cin, generator=g) * 0.5).to(device).half()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.05).to(device).half()
    bias = torch.randn(Cout, generator=g).to(device)
    cat = torch.zeros(Nimg, H, H, Ctot, device=device, dtype=torch.float16)
    out_view = cat[..., coff:coff + Cout]
    wt = gemm.pack_conv_weight(w)
    ldc = Ctot
    plan = gemm.GemmPlan(a, (Cin, H * Cin, H * H * Cin), Cin, wt, (wt.shape[1], 0, 0), wt.shape[1], wt.shape[0], Nimg, H, H, 64, 9, Cout,
                         type("P", (), {"data_ptr": lambda self: cat.data_ptr() + 2 * coff, "device": device})(), (ldc, H * ldc, H * H * ldc),
                         bias=bias, act="silu", splitk=1, block_n=block_n, cta_pair=pair)
    assert plan.can_carry_stats()
    stats = torch.zeros(Nimg, 32, 2, device=device)
    plan.add_gn_stats(stats, Ctot // 32, coff)
    plan.run()
    torch.cuda.synchronize()
    check(out_view, ref_conv(a, w, bias=bias, act="silu"), 9 * Cin)
    cpg = Ctot // 32
    x = cat.float()                                                  # channels outside [coff, coff + Cout) are zero: they add nothing
    ref_sum = x.view(Nimg, H * H, 32, cpg).sum(dim=(1, 3))
    ref_sq = (x * x).view(Nimg, H * H, 32, cpg).sum(dim=(1, 3))
    assert torch.allclose(stats[..., 0], ref_sum, rtol=2e-4, atol=2e-2), (stats[..., 0] - ref_sum).abs().max()
    assert torch.allclose(stats[..., 1], ref_sq, rtol=2e-4, atol=2e-2), (stats[..., 1] - ref_sq).abs().max()
    if coff == 0 and Ctot == Cout:
        gamma, beta = torch.randn(Cout, generator=g).to(device), torch.randn(Cout, generator=g).to(device)
        y = torch.empty_like(cat)
        _lib.call("sdf_groupnorm_apply", cat.data_ptr(), Ctot, y.data_ptr(), Ctot, Nimg, H * H, Ctot, 32, gamma.data_ptr(), beta.data_ptr(), 1e-5, 1,
                  stats.data_ptr(), _lib.stream())
        ref = F.silu(F.group_norm(x.permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
        assert (y.float() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("Nimg,Hin,Cin,Cout,pad_lo,bn,pair", [(2, 64, 320, 320, 1, 160, 0), (2, 32, 640, 640, 1, 160, 0), (2, 16, 1280, 1280, 1, 128, 0),
                                                       (1, 512, 128, 128, 0, 128, 0), (1, 256, 256, 256, 0, 256, 1), (1, 128, 512, 512, 0, 128, 0),
                                                       (2, 20, 64, 96, 1, 128, 0)])
def test_conv3x3_stride2_strided_tma(device, Nimg, Hin, Cin, Cout, pad_lo, bn, pair):
    """3x3 stride-2 convolutions read through an element-strided TMA box (sdf_gemm_plan_create_strided): the UNet's Downsample
    (pad 1 on every side, openaimodel.py:130-138) and the VAE encoder's ((0,1,0,1) zero pad then pad 0, model.py:67-79), against F.conv2d"""
    a = rnd(Nimg, Hin, Hin, Cin, device=device, seed=1)
    w = rnd(Cout, Cin, 3, 3, device=device, scale=1.0 / math.sqrt(9 * Cin), seed=2)
    bias = rnd(Cout, device=device, seed=3).float()
    Ho = Hin // 2
    out = torch.full((Nimg, Ho, Ho, Cout), float("nan"), device=device, dtype=torch.float16)
    plan = gemm.conv_plan(a.contiguous(), Cin, gemm.pack_conv_weight(w), Cout, out, taps=9, bias=bias, block_n=bn, cta_pair=pair, stride=2, pad_lo=pad_lo)
    plan.run()
    torch.cuda.synchronize()
    x = a.float().permute(0, 3, 1, 2)
    if pad_lo == 0:
        x = F.pad(x, (0, 1, 0, 1))
        ref = F.conv2d(x, w.float(), bias, stride=2, padding=0)
    else:
        ref = F.conv2d(x, w.float(), bias, stride=2, padding=1)
    check(out, ref.permute(0, 2, 3, 1), 9 * Cin)
