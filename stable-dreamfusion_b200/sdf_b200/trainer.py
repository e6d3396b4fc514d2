"""SDS training step of the -O preset: host-side mirror of the reference loop body
(nerf/utils.py:1032-1072 train_one_epoch, :439-723 train_step, :725-741 post_train_step; main.py:363-410 wiring), driving the
B200-native pieces: drop-in raymarching ops, the fused radiance field, the tcgen05 SDS engine and the fused Adan step.

Multi-GPU (new functionality, SURVEY.md §8e — the reference never initialises torch.distributed): one process per GPU, each
rank renders its own view(s) with its own RNG stream, the NeRF gradients are summed with ONE NCCL all-reduce over a flat fp32
bucket per step, and the 1/world factor rides in the fused Adan kernel's unscale.  The occupancy grid is refreshed on every rank
and rank 0's result is broadcast so marching stays identical.
"""
import math
import random

import numpy as np
import torch
import torch.nn.functional as F

from . import field, synth
from .network_grid import NeRFNetwork
from .optimizer import Adan
from .renderer import safe_normalize


def get_rays_torch(poses, focal, cx, cy, H, W):
    """nerf/utils.py:113-176 with N=-1: pixel-centre pinhole rays, unnormalised directions.  poses [B,4,4] on the device."""
    device = poses.device
    j, i = torch.meshgrid(torch.arange(H, device=device, dtype=torch.float32), torch.arange(W, device=device, dtype=torch.float32), indexing='ij')
    i = i.reshape(1, H * W) + 0.5
    j = j.reshape(1, H * W) + 0.5
    zs = -torch.ones_like(i)
    xs = -(i - cx) / focal * zs
    ys = (j - cy) / focal * zs
    directions = torch.stack((xs, ys, zs), dim=-1).expand(poses.shape[0], H * W, 3)
    # directions @ R^T as three broadcast FMAs: the 3x3 product otherwise lands on an 80 us cuBLAS gemv launch
    R = poses[:, None, :3, :3]
    rays_d = directions[..., 0:1] * R[..., 0] + directions[..., 1:2] * R[..., 1] + directions[..., 2:3] * R[..., 2]
    rays_o = poses[..., :3, 3][..., None, :].expand_as(rays_d)
    return rays_o, rays_d


class SDSTrainer:
    def __init__(self, opt, device, guidance, seed=0, rank=0, world_size=1, fused=True, prompt='a hamburger'):
        self.opt, self.device, self.guidance = opt, device, guidance
        self.rank, self.world_size = rank, world_size
        torch.manual_seed(seed)                       # identical initial parameters on every rank
        self.model = NeRFNetwork(opt, fused=fused).to(device)
        self.model.train()
        field.DIRECT_GRAD_ACCUM = True       # table / MLP gradients are scattered straight into .grad
        # main.py:368
        self.optimizer = Adan(self.model.get_params(5 * opt.lr), eps=1e-8, weight_decay=2e-5, max_grad_norm=5.0, foreach=False)
        self.optimizer.loss_scale = float(world_size)  # all-reduce SUM -> mean
        self.global_step = 0
        self.rng = np.random.default_rng(seed * 1000 + rank)       # cameras / schedule: per-rank stream
        random.seed(seed * 1000 + rank)
        torch.manual_seed(seed * 1000 + rank + 1)
        torch.cuda.manual_seed(seed * 1000 + rank + 1)
        # text embeddings (nerf/utils.py:352-377): uncond + default/front/side/back
        te = guidance.get_text_embeds
        self.embeddings = {'uncond': te(['']), 'default': te([prompt])}
        for d in ('front', 'side', 'back'):
            self.embeddings[d] = te([f'{prompt}, {d} view'])
        self._flat = None
        # multi-GPU: every rank renders 1/W of the rays of every view (balanced sample counts), guidance stays view-parallel;
        # cameras / lights / background colours then come from a stream that is identical on all ranks
        self.ray_parallel = world_size > 1 and (opt.h * opt.w) % world_size == 0
        self.rng_shared = np.random.default_rng(seed * 1000 + 999)
        # pinned staging ring for the poses: the host may run several steps ahead of the GPU, so a slot is only rewritten after
        # the copy that read it has completed (event per slot)
        n_pose = opt.batch_size * (world_size if self.ray_parallel else 1)
        self.pin_ring = [torch.zeros(n_pose, 4, 4).pin_memory() for _ in range(4)]
        self.pin_events = [None] * len(self.pin_ring)
        self.last_M = 0
        self.stage_events = None         # set to [] to record (name, cuda event) marks of the next step (bench.py --breakdown)

    def _mark(self, name):
        if self.stage_events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.stage_events.append((name, ev))

    # ------------------------------------------------------------------ data (nerf/provider.py:248-319 collate)
    def sample_views(self, n=None, rng=None):
        opt = self.opt
        B = opt.batch_size if n is None else n
        rng = self.rng if rng is None else rng
        poses, az = [], []
        for _ in range(B):
            pose, (r, th, ph) = synth.rand_pose(rng, tuple(opt.radius_range), tuple(opt.theta_range), tuple(opt.phi_range))
            poses.append(pose)
            a = ph
            if a > 180:
                a -= 360
            az.append(a)
        fov = rng.uniform(*opt.fovy_range)
        return np.stack(poses), np.array(az, np.float32), float(fov)

    def text_z(self, azimuth):
        """view-dependent prompt interpolation, nerf/utils.py:597-626"""
        z = [self.embeddings['uncond']] * len(azimuth)
        for a in azimuth:
            if -90 <= a < 90:
                r = 1 - a / 90 if a >= 0 else 1 + a / 90
                s, e = self.embeddings['front'], self.embeddings['side']
            else:
                r = 1 - (a - 90) / 90 if a >= 0 else 1 + (a + 90) / 90
                s, e = self.embeddings['side'], self.embeddings['back']
            z.append(r * s + (1 - r) * e)
        return torch.cat(z, dim=0)

    # ------------------------------------------------------------------ one optimisation step
    def train_step(self, views=None, shading=None, read_loss=False):
        opt, dev = self.opt, self.device
        self._mark('start')
        if self.global_step % opt.update_extra_interval == 0:
            self.model.update_extra_state()
            if self.world_size > 1:
                from .dist import broadcast_occupancy
                broadcast_occupancy(self.model, src=0)
        self.global_step += 1
        ray_par = self.ray_parallel and views is None
        H, W = opt.h, opt.w
        if ray_par:
            # all W * B views of the step, identical on every rank; this rank owns views [rank * B, rank * B + B)
            WS, Bv = self.world_size, opt.batch_size
            poses_np, az_all, fov = self.sample_views(WS * Bv, self.rng_shared)
            azimuth = az_all[self.rank * Bv:(self.rank + 1) * Bv]
            light_off = torch.from_numpy(self.rng_shared.standard_normal((WS * Bv, 1, 3)).astype(np.float32)).to(dev)
            bg_shared = torch.from_numpy(self.rng_shared.random(3).astype(np.float32)).to(dev)
            bg_coin = float(self.rng_shared.random())
        else:
            poses_np, azimuth, fov = self.sample_views() if views is None else views
        # host -> device: the step's only input (pinned staging)
        slot = self.global_step % len(self.pin_ring)
        if self.pin_events[slot] is not None:
            self.pin_events[slot].synchronize()
        self.pin_ring[slot][:poses_np.shape[0]].copy_(torch.from_numpy(poses_np))
        poses = self.pin_ring[slot][:poses_np.shape[0]].to(dev, non_blocking=True)
        if dev.type == 'cuda':
            self.pin_events[slot] = torch.cuda.Event()
            self.pin_events[slot].record()
        focal = H / (2 * math.tan(math.radians(fov) / 2))
        rays_o, rays_d = get_rays_torch(poses, focal, H / 2, W / 2, H, W)
        light_d = None
        if ray_par:
            # pixels rank, rank + W, ... of every view; per-view light as in nerf/renderer.py:759 (rays_o + randn(3))
            light_d = safe_normalize(rays_o[:, self.rank::WS] + light_off).reshape(-1, 3)
            rays_o = rays_o[:, self.rank::WS].reshape(1, -1, 3)
            rays_d = rays_d[:, self.rank::WS].reshape(1, -1, 3)
        B, N = (opt.batch_size, H * W) if ray_par else rays_o.shape[:2]

        # schedule (nerf/utils.py:503-535)
        exp_iter_ratio = (self.global_step - 1) / opt.iters
        if shading is None:
            if exp_iter_ratio <= opt.latent_iter_ratio:
                ambient_ratio, shading, as_latent, bg_color = 1.0, 'normal', True, None
            else:
                if exp_iter_ratio <= opt.albedo_iter_ratio:
                    ambient_ratio, shading = 1.0, 'albedo'
                else:
                    # one draw per step for the whole batch, like the reference; ray-parallel ranks must agree on it
                    u1, u2 = (float(self.rng_shared.random()), float(self.rng_shared.random())) if ray_par else (random.random(), random.random())
                    ambient_ratio = opt.min_ambient_ratio + (1.0 - opt.min_ambient_ratio) * u1
                    shading = 'textureless' if u2 >= (1.0 - opt.textureless_ratio) else 'lambertian'
                as_latent = False
                if ray_par:
                    bg_color = None if (opt.bg_radius > 0 and bg_coin > 0.5) else bg_shared
                else:
                    bg_color = None if (opt.bg_radius > 0 and random.random() > 0.5) else torch.rand(3).to(dev)
        else:
            as_latent = shading == 'latent'
            ambient_ratio = 1.0 if shading in ('albedo', 'latent') else 0.55
            bg_color = None
            shading = 'normal' if as_latent else shading

        outputs = self.model.render(rays_o, rays_d, None, H, W, staged=False, perturb=True, bg_color=bg_color, ambient_ratio=ambient_ratio,
                                    shading=shading, binarize=False, light_d=light_d)
        if ray_par:
            # rendered pixels travel to the rank that owns their view; the SDS pixel gradients come back the same way
            from .dist import exchange_pixels
            local = torch.cat([outputs['image'], outputs['weights_sum'].unsqueeze(-1)], dim=-1).reshape(WS, Bv, (H * W) // WS, 4)
            full = exchange_pixels(local, WS)                                # [B, HW, 4] complete images of my views
            chans = full if as_latent else full[..., :3]
            pred_rgb = chans.reshape(B, H, W, chans.shape[-1]).permute(0, 3, 1, 2).contiguous()
        elif as_latent:
            pred_rgb = torch.cat([outputs['image'], outputs['weights_sum'].unsqueeze(-1)], dim=-1).reshape(B, H, W, 4).permute(0, 3, 1, 2).contiguous()
        else:
            pred_rgb = outputs['image'].reshape(B, H, W, 3).permute(0, 3, 1, 2).contiguous()
        self.last_M = int(outputs['weights'].shape[0])
        self._mark('render (rays, march, field forward, composite)')

        loss = self.guidance.train_step(self.text_z(azimuth), pred_rgb, as_latent=as_latent, guidance_scale=opt.guidance_scale,
                                        grad_scale=opt.lambda_guidance)
        self._mark('guidance (VAE encode, UNet, SDS gradient, VAE data-gradient)')
        # regularisers (nerf/utils.py:686-709)
        if opt.lambda_opacity > 0:
            loss = loss + opt.lambda_opacity * (outputs['weights_sum'] ** 2).mean()
        if opt.lambda_entropy > 0:
            alphas = outputs['weights'].clamp(1e-5, 1 - 1e-5)
            loss_entropy = (-alphas * torch.log2(alphas) - (1 - alphas) * torch.log2(1 - alphas)).mean()
            loss = loss + opt.lambda_entropy * min(1, 2 * self.global_step / opt.iters) * loss_entropy
        if opt.lambda_orient > 0 and 'loss_orient' in outputs:
            loss = loss + opt.lambda_orient * outputs['loss_orient']

        self._mark('regularisers')
        loss.backward()
        self._mark('backward (composite, field backward)')
        if self.world_size > 1:
            self._allreduce_grads()
            self._mark('all-reduce')
        self.optimizer.step(zero_grad=True)      # gradients are cleared by the fused step: buffers stay allocated (and bucketed)
        self._mark('Adan step')
        if read_loss:
            return float(loss.item())          # device -> host read of the step's result (nerf/utils.py:1072)
        return loss

    def _has_grads(self):
        return any(p.grad is not None for p in self.model.parameters())

    def _allreduce_grads(self):
        """ONE NCCL all-reduce per step over the flat gradient bucket (table gradient first, then the MLPs)."""
        if self._flat is None:
            from .dist import GradBucket
            named = dict(self.model.named_parameters())
            order = [named['encoder.embeddings']] + [p for n, p in named.items() if n != 'encoder.embeddings']
            self._flat = GradBucket(order)
        self._flat.all_reduce()
