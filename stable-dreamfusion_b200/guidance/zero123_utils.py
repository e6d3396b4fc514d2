"""Drop-in `guidance.zero123_utils` (reference: guidance/zero123_utils.py:55-300): the Zero123 class with
`train_step(embeddings, pred_rgb, polar, azimuth, radius, guidance_scale=3, as_latent=False, grad_scale=1, save_guidance_path=None) -> loss`
and `get_img_embeds(x)` as nerf/utils.py:414-425,668-676 and main.py:388-390 use them (BASELINE.json config C4).

Same tcgen05 engine as guidance.sd_utils (sdf_b200/sd_engine.py), in the Zero-1-to-3 shape: 8-channel UNet input (noisy latents |
c_concat = VAE mode of the reference image), a ONE-token cross-attention context (cc_projection(CLIP image embedding | relative camera T)),
32x32 latents, VAE encoder at 256x256.  The CFG "unconditional" half is zeros for both conditionings (zero123_utils.py:170-172).  With several
reference images the UNet runs once per image and the weighted predictions are combined (zero123_utils.py:161-181) — the combination is
linear, so it is applied to both CFG halves before the fused SDS-gradient kernel.  The per-view gradient scale of
`--zero123_grad_scale angle` rides in that kernel as a device array.

Precision: the reference runs this model in fp32 ("it cannot load into fp16", zero123_utils.py:69); the engine computes in fp16 with fp32
accumulation — same tolerance as the SD path against the fp32 oracle (tests/test_gpu_zero123.py: eps 2e-2 of max, SDS gradient rel-L2 3e-2).

Weights: the reference loads `zero123-xl.ckpt` through ldm + omegaconf + pytorch_lightning and a CLIP image encoder, none of which is
available offline: the constructor takes CompVis-keyed state dicts (`weights={'unet', 'vae', 'cc_projection.weight', 'cc_projection.bias'}`)
or `weights='random'` (seeded synthetic weights of the architecture).  `get_img_embeds` needs an image encoder callable
(`.image_encoder`: [1,3,256,256] in [-1,1] -> [1,1,768]); with synthetic weights a deterministic pseudo-embedding is used, with real
weights and no encoder it raises.
"""
import hashlib
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from sdf_b200 import _lib
from sdf_b200 import sd_engine as E

from .sd_utils import unet_param_shapes, vae_param_shapes

UNET_ZERO123 = dict(E.UNET_SD15, in_channels=8)


def _angles_and_weights(polar, azimuth, radius, emb):
    """zero123_utils.py:93-152 on host floats: angle (deg) of every view to every reference view -> (angles [B][R], weights [B][R])"""
    def cart(r, th, ph):
        return (r * math.sin(th) * math.cos(ph), r * math.sin(th) * math.sin(ph), r * math.cos(th))

    def unit(v):
        n = math.sqrt(sum(c * c for c in v))
        return tuple(c / n for c in v)
    rp, ra, rr = emb['ref_polars'], emb['ref_azimuths'], emb['ref_radii']
    R = len(ra)
    refs = [unit(cart(rr[j], math.radians(rp[j]), math.radians(ra[j]))) for j in range(R)]
    angles, weights = [], []
    for p, a, r in zip(polar, azimuth, radius):
        v = unit(cart(r + rr[0], math.radians(p + rp[0]), math.radians(a + ra[0])))
        ang = [math.degrees(math.acos(max(-1.0, min(1.0, sum(x * y for x, y in zip(v, u)))))) for u in refs]
        if R > 1:
            inv = [min(1.0 / max(x, 1e-30), 100.0) for x in ang]
            m = max(inv)
            inv = [x / m for x in inv]
            inv = [0.0 if x < 0.1 else x for x in inv]
        else:
            inv = [1.0]
        ws = [w * x for w, x in zip(emb['zero123_ws'], inv)]
        m = max(ws)
        ws = [x / m for x in ws]
        ws = [0.0 if x < 0.1 else x for x in ws]
        angles.append(ang)
        weights.append(ws)
    return angles, weights


def _floats(v, n):
    if torch.is_tensor(v):
        return [float(x) for x in v.detach().reshape(-1).tolist()]
    if isinstance(v, (int, float)):
        return [float(v)] * n
    return [float(x) for x in v]


class _Zero123Loss(torch.autograd.Function):
    """loss whose gradient wrt pred_rgb is the engine's d_pred_rgb"""

    @staticmethod
    def forward(ctx, pred_rgb, z, embeddings, polar, azimuth, radius, as_latent, guidance_scale, grad_scale):
        eng = z.engine
        B, hw = pred_rgb.shape[0], eng.lat_hw
        dev = pred_rgb.device
        eng.guidance_scale, eng.grad_scale = float(guidance_scale), 1.0
        angles, ws = _angles_and_weights(polar, azimuth, radius, embeddings)
        R = len(embeddings['ref_azimuths'])
        if z.opt is not None and getattr(z.opt, 'zero123_grad_scale', 'angle') == 'None':
            scales = [1.0] * B
        else:
            scales = [min(a) / (180.0 / R) * float(grad_scale) for a in angles]
        z.view_scale.copy_(z.staging().upload(torch.tensor(scales, dtype=torch.float32)))       # pinned staging: no host sync
        # random draws in the reference's order: posterior sample (encode_imgs) -> t -> noise (zero123_utils.py:133-156)
        if not as_latent:
            eng.eps_post.copy_(torch.randn(B, 4, hw, hw, device=dev))
        t = torch.randint(z.min_step, z.max_step + 1, (B,), dtype=torch.long, device=dev)
        eng.t.copy_(t.to(torch.int32))
        eng.noise.copy_(torch.randn(B, 4, hw, hw, device=dev))
        if as_latent:
            lat = pred_rgb.detach().float()
            if lat.shape[-1] != hw or lat.shape[-2] != hw:
                lat = F.interpolate(lat, (hw, hw), mode='bilinear', align_corners=False)
            eng.latents_in.copy_(lat * 2 - 1)
            ctx.resize_from = tuple(pred_rgb.shape[-2:])
        else:
            eng.pred_rgb.copy_(pred_rgb.detach().float())
        eng.encode(as_latent)
        u = eng.unet
        if R > 1:
            z.eps_acc.zero_()
        for r in range(R):
            z.set_condition(embeddings, r, polar, azimuth, radius)
            u.runlist.run()
            if R > 1:
                w = torch.tensor([x[r] / sum(x) for x in ws] * 2, dtype=torch.float32).view(2 * B, 1, 1, 1).to(dev, non_blocking=True)
                z.eps_acc.add_(u.eps.float() * w)
        if R > 1:
            u.eps.copy_(z.eps_acc)
        eng.finish(as_latent, view_scale=z.view_scale)
        ctx.as_latent, ctx.z, ctx.in_dtype = as_latent, z, pred_rgb.dtype
        return eng.loss[0].clone()

    @staticmethod
    def backward(ctx, g):
        eng = ctx.z.engine
        if ctx.as_latent:
            d = eng.grad * (2.0 / eng.grad.shape[0])
            if tuple(d.shape[-2:]) != ctx.resize_from:
                with torch.enable_grad():
                    x = torch.zeros(d.shape[0], d.shape[1], *ctx.resize_from, device=d.device, requires_grad=True)
                    y = F.interpolate(x, d.shape[-2:], mode='bilinear', align_corners=False)
                    (d,) = torch.autograd.grad(y, x, d)
        else:
            d = eng.d_pred_rgb
        return (d * g).to(ctx.in_dtype), None, None, None, None, None, None, None, None


class Zero123(nn.Module):
    def __init__(self, device, fp16=True, config=None, ckpt=None, vram_O=False, t_range=[0.02, 0.98], opt=None, weights='random',
                 n_views=1, render_hw=64, seed=0, capture=True, unet_cfg=None, vae_cfg=None, vae_res=256):
        super().__init__()
        self.device, self.fp16, self.vram_O, self.t_range, self.opt = device, fp16, vram_O, t_range, opt
        ucfg = dict(UNET_ZERO123) if unet_cfg is None else dict(unet_cfg)
        vcfg = dict(E.VAE_SD15) if vae_cfg is None else dict(vae_cfg)
        if weights == 'random':
            unet_sd = E.random_state(unet_param_shapes(ucfg), device, seed=seed)
            vae_sd = E.random_state(vae_param_shapes(vcfg), device, seed=seed + 1)
            g = torch.Generator(device='cpu').manual_seed(seed + 2)
            cd = ucfg['context_dim']
            bound = 1.0 / math.sqrt(cd + 4)
            cc_w = (torch.rand(cd, cd + 4, generator=g) * 2 - 1) * bound
            cc_b = (torch.rand(cd, generator=g) * 2 - 1) * bound
        else:
            unet_sd, vae_sd, cc_w, cc_b = weights['unet'], weights['vae'], weights['cc_projection.weight'], weights['cc_projection.bias']
        self._synthetic = weights == 'random' or bool(isinstance(weights, dict) and weights.get('synthetic_embeddings'))
        self.image_encoder = None
        self.engine = E.SDSEngine(unet_sd, vae_sd, device, ucfg, vcfg, n_views=n_views, render_hw=render_hw, ctx_len=1, vae_res=vae_res,
                                  capture=capture)
        del unet_sd, vae_sd
        self.nv, self.cd = n_views, ucfg['context_dim']
        self.cc_w, self.cc_b = cc_w.to(device).float(), cc_b.to(device).float()
        # cc_projection as one tcgen05 plan: [B, 768 + 4 (padded to 832)] x [768, 832]^T -> the conditional context rows of the UNet
        b = E.Builder(device)
        self.cc_in = b.buf(1, 1, n_views, E._r(self.cd + 4, 64), zero=True)
        ctx_rows = self.engine.unet.ctx.view(2 * n_views, self.cd)
        ctx_rows.zero_()                                             # unconditional rows stay zero (zero123_utils.py:171)
        self.cc_out = ctx_rows[n_views:].view(1, 1, n_views, self.cd)
        b.gemm('cc_projection', E.View(self.cc_in), self.cd + 4, E._pack_linear(self.cc_w, device), self.cd, E.View(self.cc_out), bias=self.cc_b.contiguous())
        self.cc_run = b.ops[0][1]
        self.view_scale = torch.zeros(n_views, device=device)
        hw = self.engine.lat_hw
        self.eps_acc = torch.zeros(2 * n_views, hw, hw, 8, device=device)
        self._cond_key = None
        self._t_ring = None
        self.num_train_timesteps = 1000
        self.min_step = int(self.num_train_timesteps * t_range[0])
        self.max_step = int(self.num_train_timesteps * t_range[1])
        self.alphas = self.engine.acp

    # ------------------------------------------------------------------ conditioning
    @torch.no_grad()
    def get_img_embeds(self, x):
        """x [R,3,256,256] in [0,1] -> (c: R x [1,1,768] CLIP image embeddings, v: R x [1,4,32,32] VAE posterior modes)"""
        eng = self.engine
        v_eng = eng.vae
        c, v = [], []
        res = eng.vae_res
        for xx in x:
            img = xx.unsqueeze(0).to(self.device).float()
            if self.image_encoder is not None:
                c.append(self.image_encoder(img * 2 - 1).to(self.device).float().view(1, 1, self.cd))
            elif self._synthetic:
                h = int.from_bytes(hashlib.sha256(img.cpu().numpy().tobytes()).digest()[:4], 'little')
                c.append(torch.randn(1, 1, self.cd, generator=torch.Generator(device='cpu').manual_seed(h)).to(self.device))
            else:
                raise RuntimeError('Zero123 was built with real weights but no image encoder: set `.image_encoder` (CLIP ViT-L/14 image '
                                   'embedding, [1,3,256,256] in [-1,1] -> [1,1,768]); pseudo-embeddings exist only for synthetic weights')
            # VAE posterior mode = the mean half of the moments (ldm/modules/distributions/distributions.py:66-67), through the engine's encoder
            rows = []
            B = self.nv
            src = img.repeat(B, 1, 1, 1).contiguous()
            _lib.call('sdf_bilinear_forward', _lib.ptr(src), B, 3, img.shape[-2], img.shape[-1], _lib.ptr(v_eng.img), 8, res, res, 2.0, -1.0, _lib.stream())
            v_eng.fwd.run()
            v.append(v_eng.moments[:1, :, :, :4].float().permute(0, 3, 1, 2).contiguous())
        return c, v

    def staging(self):
        """pinned ring for the per-step host scalars (camera deltas, per-view scales): a pageable `copy_` would synchronise the stream"""
        if self._t_ring is None:
            self._t_ring = _lib.PinnedRing(4 * self.nv, self.device, slots=16)
        return self._t_ring

    def set_condition(self, embeddings, r, polar, azimuth, radius):
        """write reference image r's conditioning into the UNet inputs: context row = cc_projection([CLIP | T]) for every view, channels 4..7
        of the conditional half = c_concat (zero123_utils.py:161-172)"""
        B, dev = self.nv, self.device
        rp, ra, rr = embeddings['ref_polars'], embeddings['ref_azimuths'], embeddings['ref_radii']
        T = []
        for p, a, rad in zip(polar, azimuth, radius):
            pp = p + rp[0] - rp[r]
            aa = a + ra[0] - ra[r]
            if aa > 180:
                aa -= 360
            T.append([math.radians(pp), math.sin(math.radians(-aa)), math.cos(math.radians(aa)), rad + rr[0] - rr[r]])
        cin = self.cc_in.view(B, -1)
        key = (id(embeddings), r)
        if self._cond_key != key:
            cin[:, :self.cd].copy_(embeddings['c_crossattn'][r].to(dev).view(1, self.cd).expand(B, self.cd))
            cc = embeddings['c_concat'][r].to(dev).float()                      # [1,4,h,h] -> NHWC channels 4..7 of the conditional half
            self.engine.unet.x_in[B:, :, :, 4:8].copy_(cc.permute(0, 2, 3, 1).expand(B, -1, -1, -1))
            self.engine.unet.x_in[:B, :, :, 4:8].zero_()
            self._cond_key = key
        cin[:, self.cd:self.cd + 4].copy_(self.staging().upload(torch.tensor(T, dtype=torch.float32)))
        self.cc_run()

    # ------------------------------------------------------------------ SDS
    def train_step(self, embeddings, pred_rgb, polar, azimuth, radius, guidance_scale=3, as_latent=False, grad_scale=1, save_guidance_path=None):
        if save_guidance_path:
            raise NotImplementedError('guidance visualisation needs the VAE decoder, which is outside the SDS training path')
        B = pred_rgb.shape[0]
        polar, azimuth, radius = _floats(polar, B), _floats(azimuth, B), _floats(radius, B)
        return _Zero123Loss.apply(pred_rgb, self, embeddings, polar, azimuth, radius, bool(as_latent), guidance_scale, grad_scale)

    def encode_imgs(self, imgs):
        raise NotImplementedError('use train_step; the encoder runs inside the fused SDS step')

    def decode_latents(self, latents):
        raise NotImplementedError('the VAE decoder is outside the SDS hot path')

    def __call__(self, *args, **kwargs):
        raise NotImplementedError('novel-view sampling (DDIM loop + VAE decoder) is outside the SDS hot path')
