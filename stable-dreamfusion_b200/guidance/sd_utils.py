"""Drop-in `guidance.sd_utils` (reference: guidance/sd_utils.py:25-315): the StableDiffusion class with the attribute set and
`train_step(text_embeddings, pred_rgb, guidance_scale=100, as_latent=False, grad_scale=1, save_guidance_path=None) -> loss`
that nerf/utils.py:628-631 and main.py:380-382 use.  The scalar it returns back-propagates d loss / d latents = w(t)(eps_hat - eps)
through the VAE encoder exactly like the reference's MSE trick (sd_utils.py:160-161), but the whole chain — bilinear 512 resize,
VAE encode, add_noise, two UNet evaluations, CFG, SDS gradient, VAE data-gradient, resize adjoint — runs on the tcgen05 engine
(sdf_b200/sd_engine.py) instead of diffusers + cuDNN.

Weights: the reference pulls `runwayml/stable-diffusion-v1-5` through diffusers; neither is available offline, so the constructor
takes CompVis-keyed state dicts (`weights={'unet': sd, 'vae': sd}`) or `weights='random'` (seeded synthetic weights of the SD-1.5
architecture — what the benchmark uses).  A diffusers checkpoint converts key-for-key with the standard CompVis<->diffusers map.
"""
import hashlib

import torch
import torch.nn as nn
import torch.nn.functional as F

from sdf_b200 import sd_engine as E


def seed_everything(seed):
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)


def unet_param_shapes(cfg=E.UNET_SD15):
    """{CompVis key: shape} of UNetModel(use_spatial_transformer=True, transformer_depth=1, legacy=False)"""
    mc, ted, ctx = cfg['model_channels'], 4 * cfg['model_channels'], cfg['context_dim']
    s = {'time_embed.0.weight': (ted, mc), 'time_embed.0.bias': (ted,), 'time_embed.2.weight': (ted, ted), 'time_embed.2.bias': (ted,)}

    def res(p, cin, cout):
        s.update({p + '.in_layers.0.weight': (cin,), p + '.in_layers.0.bias': (cin,), p + '.in_layers.2.weight': (cout, cin, 3, 3),
                  p + '.in_layers.2.bias': (cout,), p + '.emb_layers.1.weight': (cout, ted), p + '.emb_layers.1.bias': (cout,),
                  p + '.out_layers.0.weight': (cout,), p + '.out_layers.0.bias': (cout,), p + '.out_layers.3.weight': (cout, cout, 3, 3),
                  p + '.out_layers.3.bias': (cout,)})
        if cin != cout:
            s.update({p + '.skip_connection.weight': (cout, cin, 1, 1), p + '.skip_connection.bias': (cout,)})

    def attn(p, c):
        s.update({p + '.norm.weight': (c,), p + '.norm.bias': (c,), p + '.proj_in.weight': (c, c, 1, 1), p + '.proj_in.bias': (c,),
                  p + '.proj_out.weight': (c, c, 1, 1), p + '.proj_out.bias': (c,)})
        t = p + '.transformer_blocks.0'
        for a, kd in (('attn1', c), ('attn2', ctx)):
            s.update({f'{t}.{a}.to_q.weight': (c, c), f'{t}.{a}.to_k.weight': (c, kd), f'{t}.{a}.to_v.weight': (c, kd),
                      f'{t}.{a}.to_out.0.weight': (c, c), f'{t}.{a}.to_out.0.bias': (c,)})
        s.update({t + '.ff.net.0.proj.weight': (8 * c, c), t + '.ff.net.0.proj.bias': (8 * c,), t + '.ff.net.2.weight': (c, 4 * c), t + '.ff.net.2.bias': (c,)})
        for i in (1, 2, 3):
            s.update({f'{t}.norm{i}.weight': (c,), f'{t}.norm{i}.bias': (c,)})

    def block(p, layers):
        for j, l in enumerate(layers):
            q = f'{p}.{j}'
            if l[0] == 'conv_in':
                s.update({q + '.weight': (l[2], l[1], 3, 3), q + '.bias': (l[2],)})
            elif l[0] == 'res':
                res(q, l[1], l[2])
            elif l[0] == 'attn':
                attn(q, l[1])
            elif l[0] == 'down':
                s.update({q + '.op.weight': (l[1], l[1], 3, 3), q + '.op.bias': (l[1],)})
            elif l[0] == 'up':
                s.update({q + '.conv.weight': (l[1], l[1], 3, 3), q + '.conv.bias': (l[1],)})

    inp, mid, out = E.unet_structure(cfg)
    for i, layers in enumerate(inp):
        block(f'input_blocks.{i}', layers)
    block('middle_block', mid)
    for i, (layers, _) in enumerate(out):
        block(f'output_blocks.{i}', layers)
    s.update({'out.0.weight': (mc,), 'out.0.bias': (mc,), 'out.2.weight': (cfg['out_channels'], mc, 3, 3), 'out.2.bias': (cfg['out_channels'],)})
    return s


def vae_param_shapes(cfg=E.VAE_SD15):
    ch, zc = cfg['ch'], 2 * cfg['z_channels']
    s = {'conv_in.weight': (ch, cfg['in_channels'], 3, 3), 'conv_in.bias': (ch,)}

    def res(p, cin, cout):
        s.update({p + '.norm1.weight': (cin,), p + '.norm1.bias': (cin,), p + '.conv1.weight': (cout, cin, 3, 3), p + '.conv1.bias': (cout,),
                  p + '.norm2.weight': (cout,), p + '.norm2.bias': (cout,), p + '.conv2.weight': (cout, cout, 3, 3), p + '.conv2.bias': (cout,)})
        if cin != cout:
            s.update({p + '.nin_shortcut.weight': (cout, cin, 1, 1), p + '.nin_shortcut.bias': (cout,)})

    in_mult = (1,) + tuple(cfg['ch_mult'])
    n = len(cfg['ch_mult'])
    for i in range(n):
        bin_, bout = ch * in_mult[i], ch * cfg['ch_mult'][i]
        for j in range(cfg['num_res_blocks']):
            res(f'down.{i}.block.{j}', bin_, bout)
            bin_ = bout
        if i != n - 1:
            s.update({f'down.{i}.downsample.conv.weight': (bin_, bin_, 3, 3), f'down.{i}.downsample.conv.bias': (bin_,)})
    res('mid.block_1', bin_, bin_)
    res('mid.block_2', bin_, bin_)
    s.update({'mid.attn_1.norm.weight': (bin_,), 'mid.attn_1.norm.bias': (bin_,)})
    for nm in ('q', 'k', 'v', 'proj_out'):
        s.update({f'mid.attn_1.{nm}.weight': (bin_, bin_, 1, 1), f'mid.attn_1.{nm}.bias': (bin_,)})
    s.update({'norm_out.weight': (bin_,), 'norm_out.bias': (bin_,), 'conv_out.weight': (zc, bin_, 3, 3), 'conv_out.bias': (zc,),
              'quant_conv.weight': (zc, zc, 1, 1), 'quant_conv.bias': (zc,)})
    return s


class _SDSLoss(torch.autograd.Function):
    """loss whose gradient wrt pred_rgb is the engine's d_pred_rgb (already includes grad_scale and w(t))"""

    @staticmethod
    def forward(ctx, pred_rgb, sd, as_latent, guidance_scale, grad_scale):
        eng = sd.engine
        eng.guidance_scale, eng.grad_scale = float(guidance_scale), float(grad_scale)
        B = pred_rgb.shape[0]
        hw = eng.lat_hw
        # random draws in the reference's order: posterior sample (encode_imgs) -> t -> noise (sd_utils.py:95-103)
        if not as_latent:
            eng.eps_post.copy_(torch.randn(B, 4, hw, hw, device=pred_rgb.device))
        t = torch.randint(sd.min_step, sd.max_step + 1, (B,), dtype=torch.long, device=pred_rgb.device)
        eng.t.copy_(t.to(torch.int32))
        eng.noise.copy_(torch.randn(B, 4, hw, hw, device=pred_rgb.device))
        if as_latent:
            lat = pred_rgb.detach().float()
            if lat.shape[-1] != hw or lat.shape[-2] != hw:
                lat = F.interpolate(lat, (hw, hw), mode='bilinear', align_corners=False)
            eng.latents_in.copy_(lat * 2 - 1)
            ctx.resize_from = tuple(pred_rgb.shape[-2:])
        else:
            eng.pred_rgb.copy_(pred_rgb.detach().float())
        eng.step(as_latent=as_latent)
        ctx.as_latent = as_latent
        ctx.sd = sd
        ctx.in_dtype = pred_rgb.dtype
        return eng.loss[0].clone()

    @staticmethod
    def backward(ctx, g):
        eng = ctx.sd.engine
        if ctx.as_latent:
            d = eng.grad * (2.0 / eng.grad.shape[0])          # d loss / d latents = grad / B; latents = 2 x - 1
            if tuple(d.shape[-2:]) != ctx.resize_from:
                # adjoint of the bilinear resize through autograd on a tiny tensor (latent mode with h != 64 only)
                with torch.enable_grad():            # backward() runs with grad mode off
                    x = torch.zeros(d.shape[0], d.shape[1], *ctx.resize_from, device=d.device, requires_grad=True)
                    y = F.interpolate(x, d.shape[-2:], mode='bilinear', align_corners=False)
                    (d,) = torch.autograd.grad(y, x, d)
        else:
            d = eng.d_pred_rgb
        return (d * g).to(ctx.in_dtype), None, None, None, None


class StableDiffusion(nn.Module):
    def __init__(self, device, fp16=True, vram_O=False, sd_version='1.5', hf_key=None, t_range=[0.02, 0.98], weights='random',
                 n_views=1, render_hw=64, seed=0, capture=True, synthetic_text=None):
        super().__init__()
        self.device = device
        self.sd_version = sd_version
        if sd_version != '1.5' and weights == 'random':
            raise ValueError('synthetic weights are provided for the SD-1.5 architecture')
        self.precision_t = torch.float16
        if weights == 'random':
            unet_sd = E.random_state(unet_param_shapes(), device, seed=seed)
            vae_sd = E.random_state(vae_param_shapes(), device, seed=seed + 1)
        else:
            unet_sd, vae_sd = weights['unet'], weights['vae']
        self.engine = E.SDSEngine(unet_sd, vae_sd, device, n_views=n_views, render_hw=render_hw, capture=capture)
        del unet_sd, vae_sd
        self.num_train_timesteps = 1000
        self.min_step = int(self.num_train_timesteps * t_range[0])
        self.max_step = int(self.num_train_timesteps * t_range[1])
        self.alphas = self.engine.acp
        self._text = None
        self._synthetic_text = (weights == 'random') if synthetic_text is None else bool(synthetic_text)
        self.text_encoder = None          # callable prompt list -> [n, 77, 768]; assign a CLIP text encoder when real weights are used

    @torch.no_grad()
    def get_text_embeds(self, prompt):
        """The CLIP text encoder is outside the SDS hot path (and its weights are not available offline): prompts map to
        deterministic pseudo-embeddings [len(prompt), 77, 768] so that the front/side/back interpolation of nerf/utils.py:597-626 works."""
        if self.text_encoder is not None:
            return self.text_encoder(prompt).to(self.device)
        if not self._synthetic_text:
            raise RuntimeError('StableDiffusion was built with real UNet/VAE weights but no text encoder: set `.text_encoder` to a callable '
                               '(prompt list -> [n, 77, 768] CLIP embeddings) or pass precomputed embeddings to train_step; the hash-seeded '
                               'pseudo-embeddings exist only for the synthetic-weights benchmark')
        out = []
        for p in prompt:
            h = int.from_bytes(hashlib.sha256(p.encode()).digest()[:4], 'little')
            g = torch.Generator(device='cpu').manual_seed(h)
            out.append(torch.randn(77, 768, generator=g))
        return torch.stack(out).to(self.device)

    def train_step(self, text_embeddings, pred_rgb, guidance_scale=100, as_latent=False, grad_scale=1, save_guidance_path=None):
        if save_guidance_path:
            raise NotImplementedError('guidance visualisation needs the VAE decoder, which is outside the SDS training path')
        self.engine.set_text(text_embeddings)
        return _SDSLoss.apply(pred_rgb, self, bool(as_latent), guidance_scale, grad_scale)

    def encode_imgs(self, imgs):
        raise NotImplementedError('use train_step; the encoder runs inside the fused SDS step')

    # ---- reference methods outside the SDS training path (guidance/sd_utils.py:166-300): they fail loudly instead of silently
    #      falling back to anything else
    def train_step_perpneg(self, *args, **kwargs):
        raise NotImplementedError('Perp-Neg guidance (--perpneg) batches 1 + K prompts per view; the tcgen05 engine is built for the '
                                  'CFG pair of the default path (SURVEY.md §8 scopes train_step)')

    def produce_latents(self, *args, **kwargs):
        raise NotImplementedError('text-to-image sampling needs the DDIM loop and the VAE decoder: outside the SDS hot path')

    def decode_latents(self, latents):
        raise NotImplementedError('the VAE decoder is outside the SDS hot path')

    def prompt_to_img(self, *args, **kwargs):
        raise NotImplementedError('text-to-image sampling is outside the SDS hot path')
