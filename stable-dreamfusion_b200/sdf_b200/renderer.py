"""Host-side mirror of the CUDA-ray branch of the reference's NeRFRenderer
(nerf/renderer.py:257-349 ctor + density_blob, :710-816 run_cuda, :1103-1149 update_extra_state,
:1154-1163 render dispatch).  Same method names, arguments, result keys and buffer names
(density_grid, density_bitfield, aabb_train, aabb_infer), so a checkpoint written by either loads in
the other.  Calls the drop-in `raymarching` package; the field itself comes from the subclass.

Out of scope here: DMTet, Taichi, the pure-PyTorch `run` path, mesh export.
"""
import math

import torch
import torch.nn as nn

import raymarching


def safe_normalize(x, eps=1e-20):
    # nerf/utils.py:109-110
    return x / torch.sqrt(torch.clamp(torch.sum(x * x, -1, keepdim=True), min=eps))


def custom_meshgrid(*args):
    return torch.meshgrid(*args, indexing='ij')


class NeRFRenderer(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.bound = opt.bound
        self.cascade = 1 + math.ceil(math.log2(opt.bound))
        self.grid_size = 128
        self.max_level = None
        self.dmtet = False
        self.cuda_ray = True
        self.taichi_ray = False
        self.min_near = opt.min_near
        self.density_thresh = opt.density_thresh

        aabb_train = torch.FloatTensor([-opt.bound, -opt.bound, -opt.bound, opt.bound, opt.bound, opt.bound])
        self.register_buffer('aabb_train', aabb_train)
        self.register_buffer('aabb_infer', aabb_train.clone())

        self.register_buffer('density_grid', torch.zeros([self.cascade, self.grid_size ** 3]))
        self.register_buffer('density_bitfield', torch.zeros(self.cascade * self.grid_size ** 3 // 8, dtype=torch.uint8))
        self.mean_density = 0
        self.iter_density = 0

    @torch.no_grad()
    def density_blob(self, x):
        d = (x ** 2).sum(-1)
        if self.opt.density_activation == 'exp':
            return self.opt.blob_density * torch.exp(-d / (2 * self.opt.blob_radius ** 2))
        return self.opt.blob_density * (1 - torch.sqrt(d) / self.opt.blob_radius)

    def forward(self, x, d):
        raise NotImplementedError()

    def density(self, x):
        raise NotImplementedError()

    def reset_extra_state(self):
        self.density_grid.zero_()
        self.mean_density = 0
        self.iter_density = 0

    def run_cuda(self, rays_o, rays_d, light_d=None, ambient_ratio=1.0, shading='albedo', bg_color=None, perturb=False,
                 T_thresh=1e-4, binarize=False, **kwargs):
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        device = rays_o.device

        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, self.aabb_train if self.training else self.aabb_infer)

        if light_d is None:
            light_d = safe_normalize(rays_o + torch.randn(3, device=rays_o.device))

        results = {}
        if self.training:
            xyzs, dirs, ts, rays = raymarching.march_rays_train(rays_o, rays_d, self.bound, self.density_bitfield, self.cascade,
                                                                self.grid_size, nears, fars, perturb, self.opt.dt_gamma, self.opt.max_steps)
            dirs = safe_normalize(dirs)
            if light_d.shape[0] > 1:
                flatten_rays = raymarching.flatten_rays(rays, xyzs.shape[0]).long()
                light_d = light_d[flatten_rays]
            sigmas, rgbs, normals = self(xyzs, dirs, light_d, ratio=ambient_ratio, shading=shading)
            weights, weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, ts, rays, T_thresh, binarize)

            if self.opt.lambda_orient > 0 and normals is not None:
                loss_orient = weights.detach() * (normals * dirs).sum(-1).clamp(min=0) ** 2
                results['loss_orient'] = loss_orient.mean()
            if self.opt.lambda_3d_normal_smooth > 0 and normals is not None:
                normals_perturb = self.normal(xyzs + torch.randn_like(xyzs) * 1e-2)
                results['loss_normal_perturb'] = (normals - normals_perturb).abs().mean()
            if (self.opt.lambda_2d_normal_smooth > 0 or self.opt.lambda_normal > 0) and normals is not None:
                _, _, _, normal_image = raymarching.composite_rays_train(sigmas.detach(), (normals + 1) / 2, ts, rays, T_thresh, binarize)
                results['normal_image'] = normal_image
            results['weights'] = weights
        else:
            dtype = torch.float32
            weights_sum = torch.zeros(N, dtype=dtype, device=device)
            depth = torch.zeros(N, dtype=dtype, device=device)
            image = torch.zeros(N, 3, dtype=dtype, device=device)
            n_alive = N
            rays_alive = torch.arange(n_alive, dtype=torch.int32, device=device)
            rays_t = nears.clone()
            step = 0
            while step < self.opt.max_steps:
                n_alive = rays_alive.shape[0]
                if n_alive <= 0:
                    break
                n_step = max(min(N // n_alive, 8), 1)
                xyzs, dirs, ts = raymarching.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound, self.density_bitfield,
                                                        self.cascade, self.grid_size, nears, fars, perturb if step == 0 else False,
                                                        self.opt.dt_gamma, self.opt.max_steps)
                dirs = safe_normalize(dirs)
                sigmas, rgbs, normals = self(xyzs, dirs, light_d, ratio=ambient_ratio, shading=shading)
                raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image, T_thresh, binarize)
                rays_alive = rays_alive[rays_alive >= 0]
                step += n_step

        if bg_color is None:
            if self.opt.bg_radius > 0:
                bg_color = self.background(rays_d)
            else:
                bg_color = 1
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        results['image'] = image.view(*prefix, 3)
        results['depth'] = depth.view(*prefix)
        results['weights_sum'] = weights_sum.reshape(*prefix)
        return results

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128):
        tmp_grid = -torch.ones_like(self.density_grid)
        dev = self.aabb_train.device
        X = torch.arange(self.grid_size, dtype=torch.int32, device=dev).split(S)
        Y = torch.arange(self.grid_size, dtype=torch.int32, device=dev).split(S)
        Z = torch.arange(self.grid_size, dtype=torch.int32, device=dev).split(S)
        for xs in X:
            for ys in Y:
                for zs in Z:
                    xx, yy, zz = custom_meshgrid(xs, ys, zs)
                    coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
                    indices = raymarching.morton3D(coords).long()
                    xyzs = 2 * coords.float() / (self.grid_size - 1) - 1
                    for cas in range(self.cascade):
                        bound = min(2 ** cas, self.bound)
                        half_grid_size = bound / self.grid_size
                        cas_xyzs = xyzs * (bound - half_grid_size)
                        cas_xyzs += (torch.rand_like(cas_xyzs) * 2 - 1) * half_grid_size
                        sigmas = self.density(cas_xyzs)['sigma'].reshape(-1).detach()
                        tmp_grid[cas, indices] = sigmas
        valid_mask = self.density_grid >= 0
        self.density_grid[valid_mask] = torch.maximum(self.density_grid[valid_mask] * decay, tmp_grid[valid_mask])
        self.mean_density = torch.mean(self.density_grid[valid_mask]).item()
        self.iter_density += 1
        density_thresh = min(self.mean_density, self.density_thresh)
        self.density_bitfield = raymarching.packbits(self.density_grid, density_thresh, self.density_bitfield)

    def render(self, rays_o, rays_d, mvp=None, h=None, w=None, staged=False, max_ray_batch=4096, **kwargs):
        return self.run_cuda(rays_o, rays_d, **kwargs)
