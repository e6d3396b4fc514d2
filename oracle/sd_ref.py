"""Plain-PyTorch restatement of the Stable-Diffusion-1.5-shaped guidance networks on the SDS path.

TEST INFRASTRUCTURE ONLY (oracle): the checker for the tcgen05 UNet / VAE-encoder engine and the
CPU-baseline arm of bench.py.  Never imported by the product package.

The reference calls `diffusers` (UNet2DConditionModel / AutoencoderKL / DDIMScheduler,
guidance/sd_utils.py:2,49,65,104-108,282-290), which is neither vendored nor installed (pinned only as
`diffusers >= 0.9.0`, requirements.txt:27) -> parity with real SD weights is UNPINNED.  What IS in the
reference tree is CompVis' latent-diffusion code the diffusers models were converted from; with SD-1.5
hyper-parameters it has exactly SD-1.5's parameter count.  This file restates those modules:

  UNet            ldm/modules/diffusionmodules/openaimodel.py:414-778 (UNetModel), :164-277 (ResBlock),
                  :92-120 (Upsample), :135-162 (Downsample); timestep_embedding util.py:151-171;
                  GroupNorm32 util.py:214-217 (fp32 statistics, eps 1e-5)
  transformer     ldm/modules/attention.py:221-275 (SpatialTransformer), :196-219 (BasicTransformerBlock),
                  :152-194 (CrossAttention), :37-65 (GEGLU / FeedForward); its GroupNorm uses eps 1e-6 (:75-76)
  VAE encoder     ldm/modules/diffusionmodules/model.py:368-460 (Encoder), :82-141 (ResnetBlock), :150-204 (AttnBlock),
                  :60-79 (Downsample, asymmetric (0,1,0,1) zero pad); quant_conv ldm/models/autoencoder.py:298-302,324-328;
                  posterior sample ldm/modules/distributions/distributions.py:24-37
  schedule        make_beta_schedule('linear', 1000, 0.00085, 0.012) util.py:21-25 == diffusers 'scaled_linear';
                  add_noise = sqrt(acp_t) x + sqrt(1-acp_t) eps

Module attribute names equal the CompVis state-dict keys so weights can be exchanged with the vendored
modules (tests/golden/make_golden_sd.py does that to pin this file against them).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------- UNet
def timestep_embedding(timesteps, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class GroupNorm32(nn.GroupNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


class ResBlock(nn.Module):
    def __init__(self, channels, emb_channels, out_channels):
        super().__init__()
        self.in_layers = nn.Sequential(GroupNorm32(32, channels), nn.SiLU(), nn.Conv2d(channels, out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, out_channels))
        self.out_layers = nn.Sequential(GroupNorm32(32, out_channels), nn.SiLU(), nn.Dropout(0.0),
                                        nn.Conv2d(out_channels, out_channels, 3, padding=1))
        self.skip_connection = nn.Identity() if out_channels == channels else nn.Conv2d(channels, out_channels, 1)

    def forward(self, x, emb):
        h = self.in_layers(x)
        h = h + self.emb_layers(emb).type(h.dtype)[:, :, None, None]
        h = self.out_layers(h)
        return self.skip_connection(x) + h


class CrossAttention(nn.Module):
    use_sdpa = False

    def __init__(self, query_dim, context_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        context_dim = query_dim if context_dim is None else context_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(0.0))

    def forward(self, x, context=None):
        h = self.heads
        context = x if context is None else context
        q, k, v = self.to_q(x), self.to_k(context), self.to_v(context)
        b, n, _ = q.shape
        if self.use_sdpa:        # what diffusers' default attention processor dispatches to on torch 2 (GPU reference arm only)
            heads = lambda t: t.view(b, t.shape[1], h, -1).transpose(1, 2)
            out = F.scaled_dot_product_attention(heads(q), heads(k), heads(v), scale=self.scale)
            return self.to_out(out.transpose(1, 2).reshape(b, n, -1))
        split = lambda t: t.view(b, t.shape[1], h, -1).permute(0, 2, 1, 3).reshape(b * h, t.shape[1], -1)
        q, k, v = split(q), split(k), split(v)
        sim = torch.einsum('b i d, b j d -> b i j', q, k) * self.scale
        attn = sim.softmax(dim=-1)
        out = torch.einsum('b i j, b j d -> b i d', attn, v)
        out = out.view(b, h, n, -1).permute(0, 2, 1, 3).reshape(b, n, -1)
        return self.to_out(out)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim))

    def forward(self, x):
        return self.net(x)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, context_dim):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, n_heads, d_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim, n_heads, d_head)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)

    def forward(self, x, context=None):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context=context) + x
        x = self.ff(self.norm3(x)) + x
        return x


class SpatialTransformer(nn.Module):
    def __init__(self, in_channels, n_heads, d_head, context_dim):
        super().__init__()
        inner = n_heads * d_head
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, n_heads, d_head, context_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, context=None):
        b, c, h, w = x.shape
        x_in = x
        x = self.proj_in(self.norm(x))
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, -1)
        for blk in self.transformer_blocks:
            x = blk(x, context=context)
        x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()
        return self.proj_out(x) + x_in


class Downsample(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.op = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.op(x)


class Upsample(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2, mode="nearest"))


class _Seq(nn.Sequential):
    def forward(self, x, emb, context=None):
        for layer in self:
            if isinstance(layer, ResBlock):
                x = layer(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context)
            else:
                x = layer(x)
        return x


UNET_SD15 = dict(in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, attention_resolutions=(4, 2, 1),
                 channel_mult=(1, 2, 4, 4), num_heads=8, context_dim=768)


class UNet(nn.Module):
    def __init__(self, in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, attention_resolutions=(4, 2, 1),
                 channel_mult=(1, 2, 4, 4), num_heads=8, context_dim=768):
        super().__init__()
        self.model_channels = model_channels
        ted = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))
        self.input_blocks = nn.ModuleList([_Seq(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, ted, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(SpatialTransformer(ch, num_heads, ch // num_heads, context_dim))
                self.input_blocks.append(_Seq(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(_Seq(Downsample(ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = _Seq(ResBlock(ch, ted, ch), SpatialTransformer(ch, num_heads, ch // num_heads, context_dim), ResBlock(ch, ted, ch))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [ResBlock(ch + ich, ted, model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(SpatialTransformer(ch, num_heads, ch // num_heads, context_dim))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch))
                    ds //= 2
                self.output_blocks.append(_Seq(*layers))
        self.out = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(), nn.Conv2d(model_channels, out_channels, 3, padding=1))

    def forward(self, x, timesteps, context):
        emb = self.time_embed(timestep_embedding(timesteps, self.model_channels).to(x.dtype))
        hs = []
        h = x
        for m in self.input_blocks:
            h = m(h, emb, context)
            hs.append(h)
        h = self.middle_block(h, emb, context)
        for m in self.output_blocks:
            h = torch.cat([h, hs.pop()], dim=1)
            h = m(h, emb, context)
        return self.out(h)


# ----------------------------------------------------------------------------- VAE encoder
def vae_norm(c):
    return nn.GroupNorm(32, c, eps=1e-6, affine=True)


def swish(x):
    return x * torch.sigmoid(x)


class VaeResnetBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1, self.conv1 = vae_norm(cin), nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2, self.conv2 = vae_norm(cout), nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        h = self.conv1(swish(self.norm1(x)))
        h = self.conv2(swish(self.norm2(h)))
        if hasattr(self, "nin_shortcut"):
            x = self.nin_shortcut(x)
        return x + h


class VaeAttnBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = vae_norm(c)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(c, c, 1) for _ in range(4))

    def forward(self, x):
        h_ = self.norm(x)
        q, k, v = self.q(h_), self.k(h_), self.v(h_)
        b, c, h, w = q.shape
        q = q.reshape(b, c, h * w).permute(0, 2, 1)
        k = k.reshape(b, c, h * w)
        w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
        w_ = F.softmax(w_, dim=2)
        v = v.reshape(b, c, h * w)
        h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
        return x + self.proj_out(h_)


class VaeDownsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


VAE_SD15 = dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, z_channels=4)


class VaeEncoder(nn.Module):
    """Encoder(ch, ch_mult, num_res_blocks, attn_resolutions=[], z_channels, double_z=True) + quant_conv."""

    def __init__(self, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, z_channels=4):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, ch, 3, padding=1)
        in_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        for i in range(len(ch_mult)):
            blk = nn.Module()
            bin_, bout = ch * in_mult[i], ch * ch_mult[i]
            blk.block = nn.ModuleList()
            for _ in range(num_res_blocks):
                blk.block.append(VaeResnetBlock(bin_, bout))
                bin_ = bout
            blk.attn = nn.ModuleList()
            if i != len(ch_mult) - 1:
                blk.downsample = VaeDownsample(bin_)
            self.down.append(blk)
        self.mid = nn.Module()
        self.mid.block_1 = VaeResnetBlock(bin_, bin_)
        self.mid.attn_1 = VaeAttnBlock(bin_)
        self.mid.block_2 = VaeResnetBlock(bin_, bin_)
        self.norm_out = vae_norm(bin_)
        self.conv_out = nn.Conv2d(bin_, 2 * z_channels, 3, padding=1)
        self.quant_conv = nn.Conv2d(2 * z_channels, 2 * z_channels, 1)

    def forward(self, x):
        """-> moments [B, 2z, H/8, W/8] (mean, logvar)"""
        h = self.conv_in(x)
        for i, blk in enumerate(self.down):
            for rb in blk.block:
                h = rb(h)
            if hasattr(blk, "downsample"):
                h = blk.downsample(h)
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h)))
        h = self.conv_out(swish(self.norm_out(h)))
        return self.quant_conv(h)


def posterior_sample(moments, eps):
    """DiagonalGaussianDistribution.sample with explicit noise: mean + exp(0.5*clamp(logvar,-30,20)) * eps"""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean + torch.exp(0.5 * logvar) * eps


# ----------------------------------------------------------------------------- schedule + SDS step
def alphas_cumprod(n=1000, linear_start=0.00085, linear_end=0.012):
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n, dtype=torch.float64) ** 2
    return torch.cumprod(1.0 - betas, dim=0).float()


VAE_SCALING = 0.18215


def reinit_zero_modules(model, seed=0):
    """CompVis zero-initialises ResBlock.out_layers[-1], SpatialTransformer.proj_out and UNet.out[-1]
    (openaimodel.py:230-232,720; attention.py:249): with random weights the output would be identically 0.
    Re-draw exactly those layers (Kaiming-uniform like nn.Conv2d's default) so parity tests are not vacuous."""
    g = torch.Generator().manual_seed(seed)
    targets = []
    for m in model.modules():
        if isinstance(m, ResBlock):
            targets.append(m.out_layers[-1])
        elif isinstance(m, SpatialTransformer):
            targets.append(m.proj_out)
    if isinstance(model, UNet):
        targets.append(model.out[-1])
    for conv in targets:
        fan_in = conv.weight[0].numel()
        bound = 1.0 / math.sqrt(fan_in)
        with torch.no_grad():
            conv.weight.copy_((torch.rand(conv.weight.shape, generator=g) * 2 - 1) * bound)
            conv.bias.copy_((torch.rand(conv.bias.shape, generator=g) * 2 - 1) * bound)
    return model


def sds_train_step(unet, vae, acp, text_embeddings, pred_rgb, t, noise, post_eps, guidance_scale=100.0, as_latent=False, grad_scale=1.0):
    """guidance/sd_utils.py:86-163 with the random draws (t, noise, posterior eps) passed in.
    Returns (loss, latents, grad): d loss / d latents == grad."""
    if as_latent:
        latents = F.interpolate(pred_rgb, (64, 64), mode="bilinear", align_corners=False) * 2 - 1
    else:
        rgb512 = F.interpolate(pred_rgb, (512, 512), mode="bilinear", align_corners=False)
        latents = posterior_sample(vae(2 * rgb512 - 1), post_eps) * VAE_SCALING
    with torch.no_grad():
        a = acp.to(latents.device)[t].view(-1, 1, 1, 1).to(latents.dtype)
        noisy = a.sqrt() * latents + (1 - a).sqrt() * noise
        x_in = torch.cat([noisy] * 2)
        tt = torch.cat([t] * 2)
        eps = unet(x_in, tt, text_embeddings)
        e_u, e_c = eps.chunk(2)
        eps = e_u + guidance_scale * (e_c - e_u)
        w = 1 - a
        grad = torch.nan_to_num(grad_scale * w * (eps - noise))
    targets = (latents - grad).detach()
    loss = 0.5 * F.mse_loss(latents.float(), targets.float(), reduction="sum") / latents.shape[0]
    return loss, latents, grad


# ----------------------------------------------------------------------------- Zero-1-to-3 SDS step (guidance/zero123_utils.py:113-231)
UNET_ZERO123 = dict(in_channels=8, model_channels=320, out_channels=4, num_res_blocks=2, attention_resolutions=(4, 2, 1),
                    channel_mult=(1, 2, 4, 4), num_heads=8, context_dim=768)       # zero123 config sd-objaverse-finetune-c_concat-256.yaml


def zero123_angles(polar, azimuth, radius, ref_polars, ref_azimuths, ref_radii):
    """zero123_utils.py:93-109,118-123: angle (degrees) between the novel view and every reference view; [B, R]"""
    def cart(r, th, ph):
        return torch.stack([r * torch.sin(th) * torch.cos(ph), r * torch.sin(th) * torch.sin(ph), r * torch.cos(th)], -1)
    v1 = cart(radius + ref_radii[0], torch.deg2rad(polar + ref_polars[0]), torch.deg2rad(azimuth + ref_azimuths[0]))        # [B, 3]
    v2 = cart(torch.tensor(ref_radii, dtype=torch.float32), torch.deg2rad(torch.tensor(ref_polars, dtype=torch.float32)),
              torch.deg2rad(torch.tensor(ref_azimuths, dtype=torch.float32)))                                                # [R, 3]
    v1 = v1 / v1.norm(dim=-1, keepdim=True)
    v2 = v2 / v2.norm(dim=-1, keepdim=True)
    return torch.rad2deg(torch.arccos(torch.clip(v1 @ v2.T, -1.0, 1.0)))


def zero123_weights(angles, user_ws):
    """zero123_utils.py:140-152: closeness weights per reference image; [B, R]"""
    R = angles.shape[1]
    if R > 1:
        inv = 1 / angles
        inv[inv > 100] = 100
        inv = inv / inv.max(dim=-1, keepdim=True)[0]
        inv[inv < 0.1] = 0
    else:
        inv = torch.ones(1)
    ws = torch.tensor(user_ws, dtype=torch.float32)[None, :] * inv
    ws = ws / ws.max(dim=-1, keepdim=True)[0]
    ws[ws < 0.1] = 0
    return ws


def zero123_train_step(unet, vae, cc_w, cc_b, acp, emb, pred_rgb, polar, azimuth, radius, t, noise, post_eps, guidance_scale=3.0,
                       as_latent=False, grad_scale=1.0, grad_scale_mode="angle"):
    """guidance/zero123_utils.py:113-231 with the random draws passed in.  emb: dict(c_crossattn [R x [1,1,768]], c_concat [R x [1,4,h,h]],
    ref_polars, ref_azimuths, ref_radii, zero123_ws).  polar / azimuth / radius: [B] CPU float tensors (deltas wrt the default view).
    The LatentDiffusion wrapper (ldm/models/diffusion/ddpm.py, needs pytorch_lightning: absent) is restated: apply_model = UNet on
    cat([x, c_concat], 1) with context c_crossattn; cc_projection = Linear(772, 768).  Returns (loss, latents, grad)."""
    dev = pred_rgb.device
    angles = zero123_angles(polar, azimuth, radius, emb["ref_polars"], emb["ref_azimuths"], emb["ref_radii"])
    R = len(emb["ref_azimuths"])
    if grad_scale_mode == "angle":
        gs = (angles.min(dim=1)[0] / (180 / R)) * grad_scale
    else:
        gs = torch.ones(angles.shape[0])
    gs = gs.to(dev)
    lat_hw = emb["c_concat"][0].shape[-1]
    if as_latent:
        latents = F.interpolate(pred_rgb, (lat_hw, lat_hw), mode="bilinear", align_corners=False) * 2 - 1
    else:
        rgb = F.interpolate(pred_rgb, (lat_hw * 8, lat_hw * 8), mode="bilinear", align_corners=False)
        latents = posterior_sample(vae(2 * rgb - 1), post_eps) * VAE_SCALING
    ws = zero123_weights(angles, emb["zero123_ws"]).to(dev)
    with torch.no_grad():
        a = acp.to(dev)[t].view(-1, 1, 1, 1).to(latents.dtype)
        noisy = a.sqrt() * latents + (1 - a).sqrt() * noise
        x_in = torch.cat([noisy] * 2)
        tt = torch.cat([t] * 2)
        preds = []
        for r in range(R):
            p = polar + emb["ref_polars"][0] - emb["ref_polars"][r]
            az = azimuth + emb["ref_azimuths"][0] - emb["ref_azimuths"][r]
            az = torch.where(az > 180, az - 360, az)
            rr = radius + emb["ref_radii"][0] - emb["ref_radii"][r]
            T = torch.stack([torch.deg2rad(p), torch.sin(torch.deg2rad(-az)), torch.cos(torch.deg2rad(az)), rr], dim=-1)[:, None, :].to(dev)
            cc = emb["c_crossattn"][r].to(dev)
            clip_emb = F.linear(torch.cat([cc.repeat(len(T), 1, 1), T.to(cc.dtype)], dim=-1), cc_w.to(cc.dtype), cc_b.to(cc.dtype))
            ctx = torch.cat([torch.zeros_like(clip_emb), clip_emb], dim=0)
            c_cat = emb["c_concat"][r].to(dev)
            cat = torch.cat([torch.zeros_like(c_cat).repeat(len(T), 1, 1, 1), c_cat.repeat(len(T), 1, 1, 1)], dim=0)
            eps = unet(torch.cat([x_in, cat.to(x_in.dtype)], dim=1), tt, ctx.to(x_in.dtype))
            e_u, e_c = eps.chunk(2)
            preds.append(ws[:, r][:, None, None, None] * (e_u + guidance_scale * (e_c - e_u)))
        noise_pred = torch.stack(preds).sum(dim=0) / ws.sum(dim=-1)[:, None, None, None]
        w = 1 - a
        grad = torch.nan_to_num(gs.view(-1, 1, 1, 1) * w * (noise_pred - noise))
    targets = (latents - grad).detach()
    loss = 0.5 * F.mse_loss(latents.float(), targets.float(), reduction="sum") / latents.shape[0]
    return loss, latents, grad
