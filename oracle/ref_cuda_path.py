"""Arm (A) of SURVEY.md §8d — TEST / BENCH INFRASTRUCTURE ONLY, never imported by the product.

The reference's `-O` training path as it would run on THIS GPU: its own CUDA extensions (oracle/_ref/*.so, compiled from the
unmodified sources by oracle/build_ref.py) driven with the reference's call protocol (two-pass march around a blocking
`.item()`, full-table fp32->fp16 cast per hash-grid call, zero-filled fp16 gradient table + half2 atomics, one thread per
ray in compositing), `nn.Linear` MLPs / activations / 7-point finite-difference normal as separate PyTorch ops under fp16
autocast (sdf_b200.network_grid with fused=False is that operator graph), the CompVis-shaped UNet / VAE encoder of
oracle/sd_ref.py in fp16 on cuDNN / cuBLAS with PyTorch's fused SDPA attention (what diffusers >= 0.9 on torch 2 dispatches
to), and a per-tensor (foreach=False) PyTorch Adan like the reference's optimizer.py:201-258.

What is NOT the reference here (all choices favour the reference arm): no GradScaler (its unscale pass, inf check and host
sync are skipped), no EMA, no logging / tensorboard, the host loop is sdf_b200.trainer (fewer host syncs than
nerf/utils.py:439-741).

Wrappers below are written against the pybind signatures of raymarching/src/raymarching.h:6-18, gridencoder/src/gridencoder.h:12-16
and freqencoder/src/freqencoder.h:6-9 (the same calls tests/test_gpu_raymarching.py and tests/test_gpu_encoders.py make).
"""
import importlib.util
import math
import os
import types

import numpy as np
import torch
import torch.nn.functional as F

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_cache = {}


def ext(name):
    if name not in _cache:
        path = os.path.join(_REF_DIR, name + ".so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run python oracle/build_ref.py where /root/reference exists")
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _cache[name] = mod
    return _cache[name]


# ------------------------------------------------------------------------------------------------ raymarching
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d = rays_o.contiguous().float().view(-1, 3), rays_d.contiguous().float().view(-1, 3)
    N = rays_o.shape[0]
    nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
    fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
    ext("_raymarching").near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars)
    return nears, fars


def morton3D(coords):
    N = coords.shape[0]
    out = torch.empty(N, dtype=torch.int32, device=coords.device)
    ext("_raymarching").morton3D(coords.int().contiguous(), N, out)
    return out


def packbits(grid, thresh, bitfield=None):
    grid = grid.contiguous()
    C, H3 = grid.shape
    N = C * H3 // 8
    if bitfield is None:
        bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
    ext("_raymarching").packbits(grid, N, thresh, bitfield)
    return bitfield


def flatten_rays(rays, M):
    N = rays.shape[0]
    res = torch.zeros(M, dtype=torch.int32, device=rays.device)
    ext("_raymarching").flatten_rays(rays.contiguous(), N, M, res)
    return res


def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, perturb=False, dt_gamma=0, max_steps=1024, contract=False):
    """two passes around a blocking .item(), as raymarching/raymarching.py:197-258"""
    r = ext("_raymarching")
    rays_o, rays_d = rays_o.float().contiguous().view(-1, 3), rays_d.float().contiguous().view(-1, 3)
    N = rays_o.shape[0]
    dev = rays_o.device
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    noises = torch.rand(N, dtype=rays_o.dtype, device=dev) if perturb else torch.zeros(N, dtype=rays_o.dtype, device=dev)
    rays = torch.empty(N, 2, dtype=torch.int32, device=dev)
    r.march_rays_train(rays_o, rays_d, density_bitfield, bound, contract, dt_gamma, max_steps, N, C, H, nears, fars, None, None, None, rays, counter, noises)
    M = int(counter.item())
    xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
    dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
    ts = torch.zeros(M, 2, dtype=rays_o.dtype, device=dev)
    r.march_rays_train(rays_o, rays_d, density_bitfield, bound, contract, dt_gamma, max_steps, N, C, H, nears, fars, xyzs, dirs, ts, rays, counter, noises)
    return xyzs, dirs, ts, rays


class _CompositeTrain(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, ts, rays, T_thresh, binarize):
        sigmas, rgbs = sigmas.float().contiguous(), rgbs.float().contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        weights = torch.zeros(M, dtype=sigmas.dtype, device=sigmas.device)
        ws = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
        ext("_raymarching").composite_rays_train_forward(sigmas, rgbs, ts, rays, M, N, T_thresh, binarize, weights, ws, depth, image)
        ctx.save_for_backward(sigmas, rgbs, ts, rays, ws, depth, image)
        ctx.dims = [M, N, T_thresh, binarize]
        return weights, ws, depth, image

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gw, gws, gd, gi):
        sigmas, rgbs, ts, rays, ws, depth, image = ctx.saved_tensors
        M, N, T_thresh, binarize = ctx.dims
        gs, gr = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
        ext("_raymarching").composite_rays_train_backward(gw.contiguous(), gws.contiguous(), gd.contiguous(), gi.contiguous(), sigmas, rgbs, ts, rays,
                                                          ws, depth, image, M, N, T_thresh, binarize, gs, gr)
        return gs, gr, None, None, None, None


def composite_rays_train(sigmas, rgbs, ts, rays, T_thresh=1e-4, binarize=False):
    return _CompositeTrain.apply(sigmas, rgbs, ts, rays, T_thresh, binarize)


def raymarching_namespace():
    """what sdf_b200.renderer calls on its `raymarching` module, backed by the reference extension"""
    ns = types.SimpleNamespace(near_far_from_aabb=near_far_from_aabb, morton3D=morton3D, packbits=packbits, flatten_rays=flatten_rays,
                               march_rays_train=march_rays_train, composite_rays_train=composite_rays_train)
    return ns


# ------------------------------------------------------------------------------------------------ encoders
class _GridEncode(torch.autograd.Function):
    """gridencoder/grid.py:25-96: outputs [L, B, C] permuted to [B, L*C]; backward zero-fills a half gradient table"""

    @staticmethod
    def forward(ctx, inputs, table_h, offsets, S, H, gridtype, align_corners, interp, max_level):
        inputs = inputs.float().contiguous()
        B, D = inputs.shape
        L, C = offsets.shape[0] - 1, table_h.shape[1]
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=table_h.dtype)
        ext("_gridencoder").grid_encode_forward(inputs, table_h, offsets, outputs, B, D, C, L, max_level, S, H, None, gridtype, align_corners, interp)
        ctx.save_for_backward(inputs, table_h, offsets)
        ctx.dims = [B, D, C, L, S, H, gridtype, align_corners, interp, max_level]
        return outputs.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        inputs, table_h, offsets = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, align_corners, interp, max_level = ctx.dims
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        g_table = torch.zeros_like(table_h)
        ext("_gridencoder").grid_encode_backward(grad, inputs, table_h, offsets, g_table, B, D, C, L, max_level, S, H, None, None, gridtype, align_corners, interp)
        return None, g_table, None, None, None, None, None, None, None


def patch_grid_encoder(enc):
    """enc: our drop-in GridEncoder module (parameters / offsets / level geometry are the reference's); its forward is replaced by the
    reference extension behind the reference protocol (fp32 -> fp16 table cast on EVERY call, visible to autograd)."""
    S = float(np.log2(enc.per_level_scale))

    def forward(inputs, bound=1, max_level=None):
        x = (inputs + bound) / (2 * bound)
        prefix = list(x.shape[:-1])
        x = x.view(-1, enc.input_dim)
        L = enc.offsets.shape[0] - 1
        ml = L if max_level is None else max(min(int(math.ceil(max_level * L)), L), 1)
        table_h = enc.embeddings.to(torch.half)
        out = _GridEncode.apply(x, table_h, enc.offsets, S, int(enc.base_resolution), enc.gridtype_id, enc.align_corners, enc.interp_id, ml)
        return out.view(prefix + [enc.output_dim])

    enc.forward = forward


class _FreqEncode(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, inputs, degree, output_dim):
        inputs = inputs.contiguous()
        B, D = inputs.shape
        out = torch.empty(B, output_dim, dtype=inputs.dtype, device=inputs.device)
        ext("_freqencoder").freq_encode_forward(inputs, B, D, degree, output_dim, out)
        ctx.save_for_backward(inputs, out)
        ctx.dims = [B, D, degree, output_dim]
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, out = ctx.saved_tensors
        B, D, degree, C = ctx.dims
        gi = torch.zeros_like(inputs)
        ext("_freqencoder").freq_encode_backward(grad.contiguous(), out, B, D, degree, C, gi)
        return gi, None, None


def patch_freq_encoder(enc):
    def forward(inputs, **kwargs):
        prefix = inputs.shape[:-1]
        out = _FreqEncode.apply(inputs.reshape(-1, enc.input_dim), enc.degree, enc.output_dim)
        return out.reshape(list(prefix) + [enc.output_dim])
    enc.forward = forward


# ------------------------------------------------------------------------------------------------ optimizer
class TorchAdan(torch.optim.Optimizer):
    """optimizer.py:23-258 with foreach=False: global-norm clip, then ~20 elementwise PyTorch ops per parameter tensor"""

    def __init__(self, params, lr=1e-3, betas=(0.98, 0.92, 0.99), eps=1e-8, weight_decay=0.0, max_grad_norm=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm))
        self.loss_scale = 1.0

    @torch.no_grad()
    def step(self, zero_grad=False):
        mg = self.defaults["max_grad_norm"]
        clip = torch.tensor(1.0, device=self.param_groups[0]["params"][0].device)
        if mg > 0:
            gn = torch.zeros(1, device=clip.device)
            for g in self.param_groups:
                for p in g["params"]:
                    if p.grad is not None:
                        gn.add_(p.grad.pow(2).sum())
            gn = torch.sqrt(gn)
            clip = torch.clamp(mg / (gn + self.defaults["eps"]), max=1.0)
        for group in self.param_groups:
            b1, b2, b3 = group["betas"]
            group["step"] = group.get("step", 0) + 1
            bc1, bc2, bc3s = 1.0 - b1 ** group["step"], 1.0 - b2 ** group["step"], math.sqrt(1.0 - b3 ** group["step"])
            lr, wd, eps = group["lr"], group["weight_decay"], group["eps"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                    st["exp_avg_diff"] = torch.zeros_like(p)
                grad = p.grad.mul(clip)
                if "neg_pre_grad" not in st or group["step"] == 1:
                    st["neg_pre_grad"] = grad.clone().mul_(-1.0)
                npg = st["neg_pre_grad"]
                npg.add_(grad)                                          # diff
                st["exp_avg"].mul_(b1).add_(grad, alpha=1 - b1)
                st["exp_avg_diff"].mul_(b2).add_(npg, alpha=1 - b2)
                npg.mul_(b2).add_(grad)                                  # update
                st["exp_avg_sq"].mul_(b3).addcmul_(npg, npg, value=1 - b3)
                denom = (st["exp_avg_sq"].sqrt() / bc3s).add_(eps)
                p.addcdiv_(st["exp_avg"], denom, value=-lr / bc1)
                p.addcdiv_(st["exp_avg_diff"], denom, value=-lr * b2 / bc2)
                p.div_(1 + lr * wd)
                npg.zero_().add_(grad, alpha=-1.0)
        if zero_grad:
            for g in self.param_groups:
                for p in g["params"]:
                    p.grad = None


# ------------------------------------------------------------------------------------------------ guidance
class RefGuidance(torch.nn.Module):
    """guidance/sd_utils.py:86-163 on PyTorch modules (fp16 weights, autocast) — cuDNN convolutions, cuBLAS linears, SDPA"""

    def __init__(self, device, seed=0):
        super().__init__()
        from oracle import sd_ref
        self.sd_ref = sd_ref
        self.device = device
        torch.manual_seed(seed)
        with torch.device(device):
            self.unet = sd_ref.UNet(**sd_ref.UNET_SD15)
            self.vae = sd_ref.VaeEncoder(**sd_ref.VAE_SD15)
        sd_ref.reinit_zero_modules(self.unet, seed=seed + 1)
        self.unet = self.unet.half().eval().requires_grad_(False)
        self.vae = self.vae.half().eval().requires_grad_(False)
        sd_ref.CrossAttention.use_sdpa = True
        self.acp = sd_ref.alphas_cumprod().to(device)
        self.min_step, self.max_step = 20, 980
        self.gen = torch.Generator(device=device).manual_seed(seed + 7)

    def get_text_embeds(self, prompt):
        g = torch.Generator(device="cpu").manual_seed(abs(hash(tuple(prompt))) % (2 ** 31))
        return torch.randn(len(prompt), 77, 768, generator=g).to(self.device)

    def train_step(self, text_embeddings, pred_rgb, guidance_scale=100, as_latent=False, grad_scale=1, save_guidance_path=None):
        B = pred_rgb.shape[0]
        t = torch.randint(self.min_step, self.max_step + 1, (B,), device=self.device, generator=self.gen)
        noise = torch.randn(B, 4, 64, 64, device=self.device, generator=self.gen)
        post = torch.randn(B, 4, 64, 64, device=self.device, generator=self.gen)
        with torch.autocast("cuda", dtype=torch.float16):
            loss, _, _ = self.sd_ref.sds_train_step(self.unet, self.vae, self.acp, text_embeddings.half(), pred_rgb, t, noise, post,
                                                    float(guidance_scale), bool(as_latent), float(grad_scale))
        return loss


def build_reference_trainer(opt, device, seed=0):
    """SDSTrainer whose kernels are the reference's: operator-graph network on the reference extensions, PyTorch SD, PyTorch Adan."""
    from sdf_b200 import field, renderer as R
    from sdf_b200.trainer import SDSTrainer
    # cudnn.benchmark stays off: with it the arm spends its first ~30 steps autotuning (measured 1.35 steps/s over 13 steps vs 3.19)
    guidance = RefGuidance(device, seed)
    R.raymarching = raymarching_namespace()              # the renderer's module-level `raymarching`
    tr = SDSTrainer(opt, device, guidance, seed=seed, fused=False)
    field.DIRECT_GRAD_ACCUM = False
    patch_grid_encoder(tr.model.encoder)
    if tr.model.bg_net is not None:
        patch_freq_encoder(tr.model.encoder_bg)
    tr.optimizer = TorchAdan(tr.model.get_params(5 * opt.lr), eps=1e-8, weight_decay=2e-5, max_grad_norm=5.0)
    inner = tr.train_step

    def train_step(*a, **k):
        with torch.autocast("cuda", dtype=torch.float16):       # nerf/utils.py:1052 (fp16 preset of -O)
            return inner(*a, **k)
    tr.train_step = train_step
    return tr
