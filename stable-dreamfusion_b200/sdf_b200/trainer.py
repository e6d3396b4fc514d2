"""One SDS optimisation step of the -O preset on the B200-native pieces.

What a step does is fixed by the reference (nerf/utils.py:1032-1072 loop body, :439-723 train_step, :725-741 post_train_step,
main.py:363-410 wiring, nerf/provider.py:248-319 random cameras): refresh the occupancy grid every 16 steps, draw a camera, pick the
shading / ambient ratio / background of the schedule, render, score-distillation loss + entropy / orientation regularisers, backward,
Adan step, EMA once per epoch.  How it runs here:

  pose (64 B, pinned) -> sdf_get_rays -> render_train (ONE autograd op: march with device-side sample count, fused field, composite,
  background + mix + NCHW, regularisers)  -> guidance.train_step (tcgen05 SD engine)  -> backward (engine data-gradient, one fused
  field backward scattering into the flat gradient bucket)  -> [NCCL all-reduce of the bucket]  -> fused Adan (+ fp16 table mirror,
  + EMA on epoch boundaries, + gradient zeroing)

with no host synchronisation anywhere in the step: the pose (and, ray-parallel, the per-view light offsets / background colour) go through a
pinned staging ring (_lib.PinnedRing — a pageable `tensor.to(device)` would synchronise the stream), and when the caller asks for the loss
(nerf/utils.py:1072 reads loss.item() every step) it is copied to a pinned word on a side stream as soon as it exists — before the backward —
so the host waits for that 4-byte copy, not for the step, and already enqueues the next step while backward + optimiser run.

Multi-GPU (new functionality, SURVEY.md §8e — the reference never initialises torch.distributed): one process per GPU.  Guidance is
view-parallel (each rank owns its views' UNet/VAE work) but rendering is ray-parallel — every rank renders pixels rank::W of EVERY
view from a camera stream that is identical on all ranks, two small all-to-alls carry pixels to the view's owner and pixel
gradients back — so sample counts are balanced by construction.  The NeRF gradients are summed with ONE all-reduce over the flat
bucket, launched on a side stream as soon as the backward has been enqueued; the 1/world factor rides in Adan's unscale.
"""
import math
import random

import numpy as np
import torch

from . import _lib, synth
from .ngp import InstantNGP
from .optimizer import Adan
from .render import render_train


def get_rays_torch(poses, focal, cx, cy, H, W):
    """Pinhole rays as plain tensor arithmetic (pixel centres, z = -1, unnormalised directions R (x, y, z)): the CPU-checkable
    statement of csrc/render_aux.cu:k_get_rays, pinned to the reference's get_rays by tests/test_oracle_o2_golden.py."""
    dev = poses.device
    pix = torch.arange(H * W, device=dev)
    i = (pix % W).float() + 0.5
    j = torch.div(pix, W, rounding_mode='floor').float() + 0.5
    d = torch.stack(((i - cx) / focal, -(j - cy) / focal, -torch.ones_like(i)), dim=-1)           # [HW, 3]
    R = poses[:, :3, :3]
    rays_d = d[None, :, 0:1] * R[:, None, :, 0] + d[None, :, 1:2] * R[:, None, :, 1] + d[None, :, 2:3] * R[:, None, :, 2]
    rays_o = poses[:, None, :3, 3].expand_as(rays_d)
    return rays_o, rays_d


class SDSTrainer:
    def __init__(self, opt, device, guidance, seed=0, rank=0, world_size=1, prompt='a hamburger', ema_decay=0.95, steps_per_epoch=None,
                 image_embeddings=None):
        """image_embeddings: for Zero123 guidance, the dict nerf/utils.py:414-425 builds (c_crossattn, c_concat, ref_polars, ref_azimuths,
        ref_radii, zero123_ws); text prompts are then unused."""
        self.opt, self.device, self.guidance = opt, device, guidance
        self.image_embeddings = image_embeddings
        self.rank, self.world_size = rank, world_size
        torch.manual_seed(seed)                       # identical initial parameters on every rank
        self.dmtet = bool(getattr(opt, 'dmtet', False))
        if self.dmtet:
            from .dmtet_model import DMTetNGP
            self.model = DMTetNGP(opt).to(device)
            self.model.build_lattice(device)
        else:
            self.model = InstantNGP(opt).to(device)
        self.model.train()
        # main.py:368
        self.optimizer = Adan(self.model.get_params(5 * opt.lr), eps=1e-8, weight_decay=2e-5, max_grad_norm=5.0, foreach=False)
        self.optimizer.loss_scale = float(world_size)  # all-reduce SUM -> mean
        self.model.attach_half_mirror(self.optimizer)
        # every parameter's .grad lives in ONE flat fp32 buffer from the start (all requires_grad parameters, whether or not a given
        # step touches them): the fused backward scatters into it, the all-reduce ships it, the fused Adan step clears it
        from .dist import GradBucket
        named = dict(self.model.named_parameters())
        order = [named['encoder.embeddings']] + [p for n, p in named.items() if n != 'encoder.embeddings']
        self.bucket = GradBucket(order)
        # main.py:396 ema_decay=0.95; nerf/utils.py:1090-1091 updates it once per epoch = dataset_size_train steps (main.py:116)
        self.steps_per_epoch = int(steps_per_epoch or getattr(opt, 'dataset_size_train', 100))
        if ema_decay is not None:
            self.optimizer.ema_attach(ema_decay)
        self.global_step = 0
        self.rng = np.random.default_rng(seed * 1000 + rank)       # cameras / schedule: per-rank stream
        random.seed(seed * 1000 + rank)
        torch.manual_seed(seed * 1000 + rank + 1)
        torch.cuda.manual_seed(seed * 1000 + rank + 1)
        # text embeddings (nerf/utils.py:352-377): uncond + front/side/back
        if image_embeddings is None:
            te = guidance.get_text_embeds
            self.embeddings = {'uncond': te(['']), 'default': te([prompt])}
            for d in ('front', 'side', 'back'):
                self.embeddings[d] = te([f'{prompt}, {d} view'])
        self.ray_parallel = world_size > 1 and (opt.h * opt.w) % world_size == 0 and not self.dmtet      # the mesh stage is view-parallel
        self.rng_shared = np.random.default_rng(seed * 1000 + 999)
        # pinned staging ring for the poses: the host may run several steps ahead of the GPU, so a slot is only rewritten after
        # the copy that read it has completed (event per slot)
        n_pose = opt.batch_size * (world_size if self.ray_parallel else 1) * (2 if self.dmtet else 1)          # dmtet: pose + mvp per view
        self.pin_ring = _lib.PinnedRing(n_pose * 16, device)
        self.pin_small = _lib.PinnedRing(3 * max(n_pose, 1), device, slots=8)      # light offsets, background colours
        self.comm_stream = torch.cuda.Stream(device=device) if (world_size > 1 and device.type == 'cuda') else None
        # loss read-back (nerf/utils.py:1072 reads loss.item() every step): the value exists before the backward starts, so it travels to a
        # pinned word on a side stream while backward + optimiser run, and the host waits for that 4-byte copy only — not for the step
        self.read_stream = torch.cuda.Stream(device=device)
        self.loss_dev = torch.zeros(1, device=device)
        self.loss_pin = torch.zeros(1).pin_memory()
        self.perturb = True              # march jitter (nerf/utils.py:537 perturb=True); the multi-GPU self-check turns it off
        self.stage_events = None         # set to [] to record (name, cuda event) marks of the next step (bench.py --breakdown)

    @property
    def last_M(self):
        """sample count of the most recent render (pinned mirror written by the march kernel; reading it does not synchronise)"""
        ws = getattr(self, '_last_ws', None)
        if ws is None:
            return 0
        torch.cuda.current_stream().synchronize()
        return int(ws.host_M[0])

    def _mark(self, name):
        if self.stage_events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.stage_events.append((name, ev))

    # ------------------------------------------------------------------ data
    def sample_views(self, n=None, rng=None):
        opt = self.opt
        B = opt.batch_size if n is None else n
        rng = self.rng if rng is None else rng
        poses, az, self._cam = [], [], []
        for _ in range(B):
            pose, (r, th, ph) = synth.rand_pose(rng, tuple(opt.radius_range), tuple(opt.theta_range), tuple(opt.phi_range))
            poses.append(pose)
            az.append(ph - 360 if ph > 180 else ph)
            self._cam.append((r, th, az[-1]))            # radius, polar, azimuth: the image-conditioned guidance needs them (nerf/provider.py:295-300)
        return np.stack(poses), np.array(az, np.float32), float(rng.uniform(*opt.fovy_range))

    def text_z(self, azimuth):
        """[uncond x B, view-interpolated cond x B] (nerf/utils.py:597-626: front->side for |az| < 90, side->back beyond)"""
        e = self.embeddings
        cond = []
        for a in azimuth:
            a = float(a)
            front = -90 <= a < 90
            r = 1 - abs(a) / 90 if front else 1 - (abs(a) - 90) / 90
            s, t = (e['front'], e['side']) if front else (e['side'], e['back'])
            cond.append(r * s + (1 - r) * t)
        return torch.cat([e['uncond']] * len(azimuth) + cond, dim=0)

    def _upload_poses(self, poses_np):
        return self.pin_ring.upload(np.ascontiguousarray(poses_np, dtype=np.float32))

    def _rays(self, poses, fov, first=0, stride=1):
        H, W = self.opt.h, self.opt.w
        focal = H / (2 * math.tan(math.radians(fov) / 2))
        B = poses.shape[0]
        per_view = (H * W - first + stride - 1) // stride
        rays_o = torch.empty(B * per_view, 3, device=self.device)
        rays_d = torch.empty(B * per_view, 3, device=self.device)
        _lib.call('sdf_get_rays', _lib.ptr(poses.contiguous()), B, H, W, float(focal), H / 2, W / 2, int(first), int(stride), _lib.ptr(rays_o),
                  _lib.ptr(rays_d), _lib.stream())
        return rays_o, rays_d

    def _schedule(self, shading, ray_par):
        """-> (shading, ambient_ratio, as_latent, bg_color | None) per nerf/utils.py:503-535; `shading` forces a mode (bench / tests)"""
        opt, dev = self.opt, self.device
        if shading is not None:
            as_latent = shading == 'latent'
            return ('normal' if as_latent else shading), (1.0 if shading in ('albedo', 'latent') else 0.55), as_latent, None
        ratio = (self.global_step - 1) / opt.iters
        if ratio <= opt.latent_iter_ratio:
            return 'normal', 1.0, True, None
        draw = (lambda: float(self.rng_shared.random())) if ray_par else random.random       # ray-parallel ranks must agree
        if ratio <= opt.albedo_iter_ratio:
            mode, ambient = 'albedo', 1.0
        else:
            ambient = opt.min_ambient_ratio + (1.0 - opt.min_ambient_ratio) * draw()
            mode = 'textureless' if draw() >= (1.0 - opt.textureless_ratio) else 'lambertian'
        use_net = opt.bg_radius > 0 and draw() > 0.5
        if use_net:
            bg = None
        elif ray_par:
            bg = self.pin_small.upload(self.rng_shared.random(3).astype(np.float32))
        else:
            bg = self.pin_small.upload(torch.rand(3))                 # the reference's draw (nerf/utils.py:533), staged through pinned memory
        return mode, ambient, False, bg

    # ------------------------------------------------------------------ one optimisation step
    def train_step(self, views=None, shading=None, read_loss=False):
        opt, dev = self.opt, self.device
        self._mark('start')
        if self.dmtet:
            return self._train_step_dmtet(views, shading, read_loss)
        if self.global_step % opt.update_extra_interval == 0:
            self.model.update_extra_state()
            if self.world_size > 1:
                from .dist import broadcast_occupancy
                broadcast_occupancy(self.model, src=0)
        self.global_step += 1
        self.model.entropy_ramp = min(1.0, 2 * self.global_step / opt.iters)
        ray_par = self.ray_parallel and views is None
        H, W, WS, Bv = opt.h, opt.w, self.world_size, opt.batch_size
        if ray_par:
            # all W * B views of the step, identical on every rank; this rank owns views [rank * B, rank * B + B)
            poses_np, az_all, fov = self.sample_views(WS * Bv, self.rng_shared)
            azimuth = az_all[self.rank * Bv:(self.rank + 1) * Bv]
            light_off = self.pin_small.upload(self.rng_shared.standard_normal((WS * Bv, 3)).astype(np.float32))
        else:
            poses_np, azimuth, fov = self.sample_views() if views is None else views
        poses = self._upload_poses(poses_np)                      # host -> device: the step's only input
        mode, ambient, as_latent, bg_color = self._schedule(shading, ray_par)
        if ray_par:
            rays_o, rays_d = self._rays(poses, fov, first=self.rank, stride=WS)        # pixels rank, rank + W, ... of every view
            l = poses[:, :3, 3] + light_off                                             # per-view light (nerf/renderer.py:726-727)
            l = l / torch.sqrt(torch.clamp((l * l).sum(-1, keepdim=True), min=1e-20))
            light = l.repeat_interleave((H * W) // WS, dim=0)
            self._last_shared = dict(poses=poses, fov=fov, light=l, mode=mode, ambient=ambient, as_latent=as_latent, bg_color=bg_color)
        else:
            rays_o, rays_d = self._rays(poses, fov)
            light = None
        return self.run_step(rays_o, rays_d, azimuth, mode, ambient, as_latent, bg_color, light_d=light, ray_par=ray_par, read_loss=read_loss)

    def _loss_read_begin(self, loss):
        """enqueue the device -> host copy of this step's loss behind the kernels that produced it; returns the event to wait on"""
        self.loss_dev.copy_(loss.detach().reshape(1))
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.read_stream):
            self.read_stream.wait_event(ready)
            self.loss_pin.copy_(self.loss_dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.read_stream)
        return done

    def _loss_read_end(self, done):
        done.synchronize()
        return float(self.loss_pin[0])

    def run_step(self, rays_o, rays_d, azimuth, mode, ambient, as_latent, bg_color, light_d=None, ray_par=False, read_loss=False):
        """render -> guidance -> backward -> optimiser for given rays and schedule draws (train_step draws them; parity tests replay
        the reference Trainer's).  Random numbers are consumed in the reference's order: light randn(3), march jitter rand(N),
        posterior sample, t, noise."""
        opt = self.opt
        H, W, WS, Bv = opt.h, opt.w, self.world_size, opt.batch_size
        if ray_par:
            P = (H * W) // WS
            out = render_train(self.model, rays_o, rays_d, light_d=light_d, ambient_ratio=ambient, shading=mode, bg_color=bg_color, perturb=self.perturb,
                               as_latent=as_latent, B=WS * Bv, H=1, W=P, direct_grads=True)
            from .dist import exchange_pixels
            C = out['pred_rgb'].shape[1]
            local = out['pred_rgb'].reshape(WS, Bv, C, P).permute(0, 1, 3, 2)          # [owner rank, view, pixel, channel]
            full = exchange_pixels(local, WS)                                          # [Bv, HW, C] complete images of my views
            pred_rgb = full.reshape(Bv, H, W, C).permute(0, 3, 1, 2).contiguous()
        else:
            B = rays_o.reshape(-1, 3).shape[0] // (H * W)
            out = render_train(self.model, rays_o, rays_d, light_d=light_d, ambient_ratio=ambient, shading=mode, bg_color=bg_color, perturb=self.perturb,
                               as_latent=as_latent, B=B, H=H, W=W, direct_grads=True)
            pred_rgb = out['pred_rgb']
        self._last_ws = self.model.workspace(rays_o.reshape(-1, 3).shape[0])
        self.last_pred_rgb = pred_rgb
        self._mark('render (rays, march, field forward, composite, background, regularisers)')

        if self.image_embeddings is not None:
            # Zero-1-to-3 guidance (nerf/utils.py:668-676): camera deltas wrt the default view instead of a text embedding
            cam = self._cam[self.rank * Bv:(self.rank + 1) * Bv] if ray_par else self._cam
            polar = [c[1] - getattr(opt, 'default_polar', 90.0) for c in cam]
            azim = [c[2] - getattr(opt, 'default_azimuth', 0.0) for c in cam]
            rad = [c[0] - getattr(opt, 'default_radius', 3.2) for c in cam]
            loss = self.guidance.train_step(self.image_embeddings, pred_rgb, polar, azim, rad, guidance_scale=opt.guidance_scale,
                                            as_latent=as_latent, grad_scale=opt.lambda_guidance)
        else:
            loss = self.guidance.train_step(self.text_z(azimuth), pred_rgb, as_latent=as_latent, guidance_scale=opt.guidance_scale,
                                            grad_scale=opt.lambda_guidance)
        self._mark('guidance (VAE encode, UNet, SDS gradient, VAE data-gradient)')
        loss = loss + out['reg']                                    # entropy + orientation (nerf/utils.py:690-704), lambdas folded in
        if opt.lambda_opacity > 0:
            loss = loss + opt.lambda_opacity * (out['weights_sum'] ** 2).mean()
        pending = self._loss_read_begin(loss) if read_loss else None
        loss.backward()
        self._mark('backward (background, composite, field backward)')
        if self.world_size > 1:
            self._allreduce_grads()
            self._mark('all-reduce')
        # gradients are cleared by the fused step (buffers stay allocated inside the bucket); EMA rides along on epoch boundaries
        self.optimizer.step(zero_grad=True, ema=(self.global_step % self.steps_per_epoch == 0))
        self._mark('Adan step')
        if read_loss:
            return self._loss_read_end(pending)          # device -> host read of the step's result (nerf/utils.py:1072)
        return loss

    # ------------------------------------------------------------------ DMTet stage (BASELINE config C5)
    def _train_step_dmtet(self, views, shading, read_loss):
        """one fine-tuning step of the mesh stage (nerf/utils.py:439-723 with opt.dmtet, nerf/renderer.py:862-954): sample a view, extract
        + rasterise + texture + shade the mesh at opt.h x opt.w, SDS through the same guidance engine, mesh regularisers, fused Adan over
        (table, MLP, background net, sdf, deform).  One view per GPU; gradients are all-reduced like the volume stage's."""
        opt, dev = self.opt, self.device
        assert opt.batch_size == 1, 'the DMTet stage renders one view per step and GPU'
        if self.global_step % opt.update_extra_interval == 0:
            self.model.update_extra_state()                  # the reference keeps refreshing the grid while cuda_ray is set (nerf/utils.py:1029-1031)
        self.global_step += 1
        poses_np, azimuth, fov = self.sample_views() if views is None else views
        H, W = opt.h, opt.w
        focal = H / (2 * math.tan(math.radians(fov) / 2))
        near, far = float(opt.min_near), 1000.0
        proj = np.array([[2 * focal / W, 0, 0, 0], [0, -2 * focal / H, 0, 0], [0, 0, -(far + near) / (far - near), -(2 * far * near) / (far - near)],
                         [0, 0, -1, 0]], dtype=np.float32)                       # nerf/provider.py:222-229
        mvp_np = (proj @ np.linalg.inv(poses_np[0].astype(np.float64)).astype(np.float32))[None]
        both = self._upload_poses(np.concatenate([poses_np[:1], mvp_np], 0).astype(np.float32))      # host -> device: 128 bytes
        poses, mvp = both[:1], both[1]
        mode, ambient, _, bg_color = self._schedule(shading if shading != 'latent' else 'albedo', False)
        if shading is None and mode == 'normal':             # no latent warm-up in this stage (main.py:270-272); step 1 would draw it
            mode, ambient = 'lambertian', 1.0
        rays_o, rays_d = self._rays(poses, fov)
        out = self.model.render_mesh(mvp, rays_d, poses[0, :3, 3], H, W, ambient_ratio=ambient, shading=mode, bg_color=bg_color,
                                     antialias=getattr(opt, 'mesh_antialias', True))
        pred_rgb = out['pred_rgb']
        self.last_pred_rgb = pred_rgb
        self._mark('mesh render (marching tets, normals, rasterise, texture, shade, antialias, background)')
        loss = self.guidance.train_step(self.text_z(azimuth), pred_rgb, as_latent=False, guidance_scale=opt.guidance_scale, grad_scale=opt.lambda_guidance)
        self._mark('guidance (VAE encode, UNet, SDS gradient, VAE data-gradient)')
        if 'normal_loss' in out:                              # nerf/utils.py:712-717
            loss = loss + opt.lambda_mesh_normal * out['normal_loss'] + opt.lambda_mesh_laplacian * out['lap_loss']
        pending = self._loss_read_begin(loss) if read_loss else None
        loss.backward()
        self._mark('backward (shade, rasterise, normals, regularisers, marching tets, field backward)')
        if self.world_size > 1:
            self._allreduce_grads()
            self._mark('all-reduce')
        self.optimizer.step(zero_grad=True, ema=(self.global_step % self.steps_per_epoch == 0))
        self._mark('Adan step')
        if read_loss:
            return self._loss_read_end(pending)
        return loss

    def _allreduce_grads(self):
        """ONE NCCL all-reduce per step over the flat gradient bucket, on a side stream: it starts as soon as the backward kernels
        have run and the optimiser step waits on it on the device, so the host keeps enqueueing."""
        if self.comm_stream is None:
            self.bucket.all_reduce()
            return
        self.comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            self.bucket.all_reduce()
        torch.cuda.current_stream().wait_stream(self.comm_stream)
