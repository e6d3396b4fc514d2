"""Drop-in `raymarching` package: same public names, positional order, defaults and
return tuples as the reference's raymarching/raymarching.py (:61,92,116,138,167,191,
258,317,371,398), backed by libsdf_b200.so (csrc/raymarch.cu) through the C ABI.

Differences from the reference wrappers that callers cannot observe:
  * kernels run on torch's current stream (the reference uses the legacy default stream);
  * sample offsets in `rays[:,0]` are the exclusive prefix sum in ray order
    (deterministic) instead of atomicAdd order;
  * no CPU tensors are silently moved to the GPU twice; inputs are moved once, like the reference.
"""
import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from sdf_b200 import _lib

_cfwd32 = custom_fwd(device_type='cuda', cast_inputs=torch.float32)
_cbwd = custom_bwd(device_type='cuda')


def _cuda(t):
    return t if t.is_cuda else t.cuda()


class _near_far_from_aabb(Function):
    @staticmethod
    @_cfwd32
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        rays_o = _cuda(rays_o).contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).contiguous().view(-1, 3)
        aabb = _cuda(aabb).float().contiguous()
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        _lib.call('sdf_near_far_from_aabb', _lib.ptr(rays_o), _lib.ptr(rays_d), _lib.ptr(aabb), N, min_near,
                  _lib.ptr(nears), _lib.ptr(fars), _lib.stream())
        return nears, fars


near_far_from_aabb = _near_far_from_aabb.apply


class _sph_from_ray(Function):
    @staticmethod
    @_cfwd32
    def forward(ctx, rays_o, rays_d, radius):
        rays_o = _cuda(rays_o).contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).contiguous().view(-1, 3)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
        _lib.call('sdf_sph_from_ray', _lib.ptr(rays_o), _lib.ptr(rays_d), radius, N, _lib.ptr(coords), _lib.stream())
        return coords


sph_from_ray = _sph_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        coords = _cuda(coords).int().contiguous()
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        _lib.call('sdf_morton3D', _lib.ptr(coords), N, _lib.ptr(indices), _lib.stream())
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        indices = _cuda(indices).int().contiguous()
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        _lib.call('sdf_morton3D_invert', _lib.ptr(indices), N, _lib.ptr(coords), _lib.stream())
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    @_cfwd32
    def forward(ctx, grid, thresh, bitfield=None):
        grid = _cuda(grid).contiguous()
        C = grid.shape[0]
        H3 = grid.shape[1]
        N = C * H3 // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        _lib.call('sdf_packbits', _lib.ptr(grid), N, float(thresh), _lib.ptr(bitfield), _lib.stream())
        return bitfield


packbits = _packbits.apply


class _flatten_rays(Function):
    @staticmethod
    def forward(ctx, rays, M):
        rays = _cuda(rays).contiguous()
        N = rays.shape[0]
        res = torch.zeros(M, dtype=torch.int, device=rays.device)
        _lib.call('sdf_flatten_rays', _lib.ptr(rays), N, M, _lib.ptr(res), _lib.stream())
        return res


flatten_rays = _flatten_rays.apply


class _march_rays_train(Function):
    @staticmethod
    @_cfwd32
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, perturb=False, dt_gamma=0, max_steps=1024, contract=False):
        rays_o = _cuda(rays_o).float().contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).float().contiguous().view(-1, 3)
        density_bitfield = _cuda(density_bitfield).contiguous()
        nears = nears.contiguous()
        fars = fars.contiguous()
        N = rays_o.shape[0]
        dev = rays_o.device

        counter = torch.empty(1, dtype=torch.int32, device=dev)
        noises = torch.rand(N, dtype=rays_o.dtype, device=dev) if perturb else None
        rays = torch.empty(N, 2, dtype=torch.int32, device=dev)
        args = (_lib.ptr(rays_o), _lib.ptr(rays_d), _lib.ptr(density_bitfield), float(bound), int(bool(contract)), float(dt_gamma),
                int(max_steps), N, int(C), int(H), _lib.ptr(nears), _lib.ptr(fars), _lib.ptr(noises))
        _lib.call('sdf_march_rays_train_count', *args, _lib.ptr(rays), _lib.ptr(counter), None, _lib.stream())

        # the (xyzs, dirs, ts) return contract needs M on the host (reference: raymarching.py:245)
        M = int(counter.item())

        xyzs = torch.empty(M, 3, dtype=rays_o.dtype, device=dev)
        dirs = torch.empty(M, 3, dtype=rays_o.dtype, device=dev)
        ts = torch.empty(M, 2, dtype=rays_o.dtype, device=dev)
        _lib.call('sdf_march_rays_train_write', *args, _lib.ptr(xyzs), _lib.ptr(dirs), _lib.ptr(ts), _lib.ptr(rays), M, _lib.stream())
        return xyzs, dirs, ts, rays


march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    @_cfwd32
    def forward(ctx, sigmas, rgbs, ts, rays, T_thresh=1e-4, binarize=False):
        sigmas = sigmas.float().contiguous()
        rgbs = rgbs.float().contiguous()
        ts = ts.contiguous()
        rays = rays.contiguous()
        M = sigmas.shape[0]
        N = rays.shape[0]
        weights = torch.zeros(M, dtype=sigmas.dtype, device=sigmas.device)   # rays with offset+count > M leave their slots untouched
        weights_sum = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
        _lib.call('sdf_composite_rays_train_forward', _lib.ptr(sigmas), _lib.ptr(rgbs), _lib.ptr(ts), _lib.ptr(rays), M, N,
                  float(T_thresh), int(bool(binarize)), _lib.ptr(weights), _lib.ptr(weights_sum), _lib.ptr(depth), _lib.ptr(image),
                  _lib.stream())
        ctx.save_for_backward(sigmas, rgbs, ts, rays, weights_sum, depth, image)
        ctx.dims = [M, N, T_thresh, binarize]
        return weights, weights_sum, depth, image

    @staticmethod
    @_cbwd
    def backward(ctx, grad_weights, grad_weights_sum, grad_depth, grad_image):
        grad_weights = grad_weights.contiguous()
        grad_weights_sum = grad_weights_sum.contiguous()
        grad_depth = grad_depth.contiguous()
        grad_image = grad_image.contiguous()
        sigmas, rgbs, ts, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh, binarize = ctx.dims
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        _lib.call('sdf_composite_rays_train_backward', _lib.ptr(grad_weights), _lib.ptr(grad_weights_sum), _lib.ptr(grad_depth),
                  _lib.ptr(grad_image), _lib.ptr(sigmas), _lib.ptr(rgbs), _lib.ptr(ts), _lib.ptr(rays), _lib.ptr(weights_sum),
                  _lib.ptr(depth), _lib.ptr(image), M, N, float(T_thresh), int(bool(binarize)), _lib.ptr(grad_sigmas),
                  _lib.ptr(grad_rgbs), _lib.stream())
        return grad_sigmas, grad_rgbs, None, None, None, None


composite_rays_train = _composite_rays_train.apply


class _march_rays(Function):
    @staticmethod
    @_cfwd32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, perturb=False, dt_gamma=0, max_steps=1024, contract=False):
        rays_o = _cuda(rays_o).float().contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).float().contiguous().view(-1, 3)
        M = n_alive * n_step
        dev = rays_o.device
        # zero-initialised: composite_rays reads ts[...,0] == 0 as the end-of-ray sentinel
        xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
        dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
        ts = torch.zeros(M, 2, dtype=rays_o.dtype, device=dev)
        noises = torch.rand(n_alive, dtype=rays_o.dtype, device=dev) if perturb else None
        _lib.call('sdf_march_rays', int(n_alive), int(n_step), _lib.ptr(rays_alive), _lib.ptr(rays_t), _lib.ptr(rays_o), _lib.ptr(rays_d),
                  float(bound), int(bool(contract)), float(dt_gamma), int(max_steps), int(C), int(H), _lib.ptr(density_bitfield),
                  _lib.ptr(near), _lib.ptr(far), _lib.ptr(xyzs), _lib.ptr(dirs), _lib.ptr(ts), _lib.ptr(noises), _lib.stream())
        return xyzs, dirs, ts


march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    @_cfwd32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image, T_thresh=1e-2, binarize=False):
        sigmas = sigmas.float().contiguous()
        rgbs = rgbs.float().contiguous()
        _lib.call('sdf_composite_rays', int(n_alive), int(n_step), float(T_thresh), int(bool(binarize)), _lib.ptr(rays_alive),
                  _lib.ptr(rays_t), _lib.ptr(sigmas), _lib.ptr(rgbs), _lib.ptr(ts), _lib.ptr(weights_sum), _lib.ptr(depth),
                  _lib.ptr(image), _lib.stream())
        return tuple()


composite_rays = _composite_rays.apply
