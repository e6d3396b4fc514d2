"""The -O backbone as this framework holds it: hash table + two tiny MLPs + occupancy grid, evaluated only through fused kernels.

State-dict keys match the reference's NeRFNetwork (nerf/network_grid.py:43-66, nerf/renderer.py:285-300) — encoder.embeddings,
encoder.offsets, sigma_net.net.{0,1,2}.{weight,bias}, bg_net.net.{0,1}.{weight,bias}, aabb_train, aabb_infer, density_grid,
density_bitfield — so checkpoints move both ways; the method names a renderer or trainer calls on the reference class (forward,
density, normal, background, render, update_extra_state, get_params) exist with the same arguments and results.  Nothing here is
an operator graph: forward/density/normal are one launch of csrc/fused_field.cu, render() in training mode is sdf_b200.render
(device-side sample count, no host sync), render() in eval mode is sdf_b200.render_eval (on-device alive-ray compaction), and
update_extra_state() is four launches per cascade with the density threshold read on the device.  The operator-by-operator graph of
the reference is NOT re-typed here: tests run the reference's own nerf/network_grid.py on the drop-in ops (oracle/ref_harness.py).
"""
import math

import numpy as np
import torch
import torch.nn as nn

from gridencoder import GridEncoder

from . import _lib
from .field import fused_field


class _LinearStack(nn.Module):
    """parameter holder with the reference MLP's key layout (`net.{i}.weight|bias`); never called as a module"""

    def __init__(self, dims):
        super().__init__()
        self.net = nn.ModuleList([nn.Linear(a, b, bias=True) for a, b in zip(dims[:-1], dims[1:])])


class InstantNGP(nn.Module):
    def __init__(self, opt):
        super().__init__()
        if opt.density_activation != 'exp':
            raise NotImplementedError('the fused field implements the exp density activation of the -O preset')
        self.opt = opt
        self.bound = float(opt.bound)
        self.cascade = 1 + math.ceil(math.log2(opt.bound))
        self.grid_size = 128
        self.max_level = None
        self.cuda_ray, self.dmtet, self.taichi_ray = True, False, False
        self.half_round = bool(getattr(opt, 'fp16', True))       # round the background net where fp16 autocast rounds
        self.encoder = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                                   desired_resolution=2048 * opt.bound, gridtype='hash', align_corners=False, interpolation='smoothstep')
        self.sigma_net = _LinearStack([self.encoder.output_dim, 64, 64, 4])
        self.bg_net = _LinearStack([3 + 3 * 2 * 6, 32, 3]) if opt.bg_radius > 0 else None
        box = torch.tensor([-opt.bound] * 3 + [opt.bound] * 3, dtype=torch.float32)
        self.register_buffer('aabb_train', box)
        self.register_buffer('aabb_infer', box.clone())
        self.register_buffer('density_grid', torch.zeros(self.cascade, self.grid_size ** 3))
        self.register_buffer('density_bitfield', torch.zeros(self.cascade * self.grid_size ** 3 // 8, dtype=torch.uint8))
        self.register_buffer('_occ_acc', torch.zeros(3), persistent=False)          # (sum, count, mean) of the last refresh
        self.iter_density = 0
        self.entropy_ramp = 1.0                 # min(1, 2 * step / iters), set by the trainer (nerf/utils.py:693)
        self._mirror = None                     # fp16 working copy of the table, kept current by the fused Adan step
        self._mirror_valid = False
        self._ws = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_mirror())

    # ------------------------------------------------------------------ parameters
    def get_params(self, lr):
        groups = [{'params': [self.encoder.embeddings], 'lr': lr * 10}, {'params': list(self.sigma_net.parameters()), 'lr': lr}]
        if self.bg_net is not None:
            groups.append({'params': list(self.bg_net.parameters()), 'lr': lr})
        return groups

    def invalidate_mirror(self):
        self._mirror_valid = False

    def attach_half_mirror(self, optimizer):
        """let the fused optimiser write the fp16 table next to the fp32 update (no per-step 73 MB re-cast)"""
        emb = self.encoder.embeddings
        self._mirror = emb.detach().to(torch.half).contiguous()
        optimizer.half_mirrors[id(emb)] = self._mirror
        self._mirror_valid = True

    def table_half(self):
        emb = self.encoder.embeddings
        if self._mirror_valid and self._mirror is not None and self._mirror.device == emb.device:
            return self._mirror
        return emb.detach().to(torch.half)

    @property
    def mean_density(self):
        return float(self._occ_acc[2].item())

    def field_cfg(self):
        e = self.encoder
        L = e.num_levels
        active = L if self.max_level is None else max(min(int(math.ceil(self.max_level * L)), L), 1)
        return dict(offsets=e.offsets, L=L, levels_active=active, S=float(np.log2(e.per_level_scale)), H=e.base_resolution,
                    smoothstep=e.interp_id == 1, blob_density=float(self.opt.blob_density), blob_radius=float(self.opt.blob_radius))

    def workspace(self, N):
        from .render import RenderWorkspace
        key = (int(N), self.encoder.embeddings.device)
        if key not in self._ws:
            self._ws[key] = RenderWorkspace(int(N), int(self.opt.max_steps), key[1])
        return self._ws[key]

    # ------------------------------------------------------------------ point queries (one fused launch each)
    def _field(self, x, light, ratio, shading, want_color=True):
        n = self.sigma_net.net
        c = self.field_cfg()
        return fused_field(x, self.encoder.embeddings, n[0].weight, n[0].bias, n[1].weight, n[1].bias, n[2].weight, n[2].bias, c['offsets'], light,
                           shading=shading, ratio=ratio, bound=self.bound, per_level_scale=self.encoder.per_level_scale, base_resolution=c['H'],
                           smoothstep=c['smoothstep'], levels_active=c['levels_active'], blob_density=c['blob_density'],
                           blob_radius=c['blob_radius'], want_color=want_color, table_half=self.table_half() if self._mirror_valid else None)

    def forward(self, x, d, l=None, ratio=1, shading='albedo'):
        """-> sigma [M], color [M,3], normal [M,3] | None   (nerf/network_grid.py:104-130; d is unused by this backbone)"""
        return self._field(x, l, ratio, shading)

    def density(self, x):
        sigma, albedo, _ = self._field(x, None, 1.0, 'albedo')
        return {'sigma': sigma, 'albedo': albedo}

    def normal(self, x):
        return self._field(x, torch.zeros(3, device=x.device), 1.0, 'normal')[2]

    def background(self, d):
        """sigmoid(bg_net(freq_encode(d))) [.., 3] through the fused background kernel (inference helper; no autograd)"""
        d2 = d.detach().float().contiguous().view(-1, 3)
        N = d2.shape[0]
        bn = self.bg_net.net
        zero3, zero1 = torch.zeros(N, 3, device=d2.device), torch.zeros(N, device=d2.device)
        bg = torch.empty(N, 3, device=d2.device)
        img = torch.empty(N, 3, device=d2.device)
        _lib.call('sdf_background_forward', _lib.ptr(d2), N, _lib.ptr(bn[0].weight), _lib.ptr(bn[0].bias), _lib.ptr(bn[1].weight), _lib.ptr(bn[1].bias),
                  None, int(self.half_round), _lib.ptr(zero3), _lib.ptr(zero1), _lib.ptr(bg), _lib.ptr(img), None, 0, 3, _lib.stream())
        return bg.view(*d.shape[:-1], 3)

    # ------------------------------------------------------------------ occupancy grid
    def reset_extra_state(self):
        self.density_grid.zero_()
        self._occ_acc.zero_()
        self.iter_density = 0

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128):
        """nerf/renderer.py:1103-1149 on the device: per cascade — jittered cell points (Morton order), one density-only field launch,
        decayed max-update + running mean; then bit packing against min(mean, density_thresh) with the mean read from device memory."""
        G, dev = self.grid_size, self.density_grid.device
        n = G ** 3
        st = _lib.stream()
        self._occ_acc.zero_()
        sn = self.sigma_net.net
        c = self.field_cfg()
        table = self.table_half()
        xyz = torch.empty(n, 3, device=dev)
        sig = torch.empty(n, device=dev)
        for cas in range(self.cascade):
            b = min(2 ** cas, self.bound)
            noise = torch.rand(n, 3, device=dev)
            _lib.call('sdf_occupancy_points', _lib.ptr(noise), n, G, float(b), _lib.ptr(xyz), st)
            _lib.call('sdf_field_forward', _lib.ptr(xyz), n, None, _lib.ptr(table), _lib.ptr(c['offsets']), c['L'], c['levels_active'], c['S'], int(c['H']),
                      int(c['smoothstep']), *[_lib.ptr(t) for t in (sn[0].weight, sn[0].bias, sn[1].weight, sn[1].bias, sn[2].weight, sn[2].bias)],
                      self.bound, c['blob_density'], c['blob_radius'], 0, None, 0, 1.0, _lib.ptr(sig), None, None, None, None, st)
            _lib.call('sdf_occupancy_update', _lib.ptr(self.density_grid[cas]), _lib.ptr(sig), n, float(decay), _lib.ptr(self._occ_acc), st)
        _lib.call('sdf_packbits_mean', _lib.ptr(self.density_grid), self.cascade * n // 8, _lib.ptr(self._occ_acc), float(self.opt.density_thresh),
                  _lib.ptr(self.density_bitfield), _lib.ptr(self._occ_acc[2:]), st)
        self.iter_density += 1

    # ------------------------------------------------------------------ rendering
    def render(self, rays_o, rays_d, mvp=None, h=None, w=None, staged=False, max_ray_batch=4096, **kwargs):
        """same call as NeRFRenderer.render (nerf/renderer.py:1154-1163) -> dict(image, depth, weights_sum[, weights, loss_orient])"""
        prefix = rays_o.shape[:-1]
        B = rays_o.shape[0] if rays_o.dim() == 3 else 1
        if self.training:
            from .render import render_train
            N = rays_o.reshape(-1, 3).shape[0]
            H = h if h is not None else int(round(math.sqrt(N // B)))
            W = w if w is not None else (N // B) // H
            out = render_train(self, rays_o, rays_d, light_d=kwargs.get('light_d'), ambient_ratio=kwargs.get('ambient_ratio', 1.0),
                               shading=kwargs.get('shading', 'albedo'), bg_color=kwargs.get('bg_color'), perturb=kwargs.get('perturb', False),
                               T_thresh=kwargs.get('T_thresh', 1e-4), B=B, H=H, W=W)
        else:
            from .render_eval import render_eval
            out = render_eval(self, rays_o, rays_d, light_d=kwargs.get('light_d'), ambient_ratio=kwargs.get('ambient_ratio', 1.0),
                              shading=kwargs.get('shading', 'albedo'), bg_color=kwargs.get('bg_color'), perturb=kwargs.get('perturb', False),
                              T_thresh=kwargs.get('T_thresh', 1e-4))
        out['image'] = out['image'].view(*prefix, 3)
        out['depth'] = out['depth'].view(*prefix)
        out['weights_sum'] = out['weights_sum'].view(*prefix)
        return out
