"""numpy front-end of oracle/sdf_oracle.c (CPU restatement of the reference's
native operators).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg; never from the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsdf_oracle.so")
_SRC = os.path.join(_HERE, "sdf_oracle.c")


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_march_rays_train.restype = C.c_uint32
        _lib.oracle_grid_resolution.restype = C.c_uint32
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3), _f32(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().oracle_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), C.c_uint32(N), C.c_float(min_near), _p(nears), _p(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    coords = np.empty((N, 2), np.float32)
    lib().oracle_sph_from_ray(_p(rays_o), _p(rays_d), C.c_float(radius), C.c_uint32(N), _p(coords))
    return coords


def morton3D(coords):
    coords = _i32(coords)
    N = coords.shape[0]
    out = np.empty(N, np.int32)
    lib().oracle_morton3D(_p(coords), C.c_uint32(N), _p(out))
    return out


def morton3D_invert(indices):
    indices = _i32(indices)
    N = indices.shape[0]
    out = np.empty((N, 3), np.int32)
    lib().oracle_morton3D_invert(_p(indices), C.c_uint32(N), _p(out))
    return out


def packbits(grid, thresh):
    grid = _f32(grid)
    N = grid.size // 8
    out = np.empty(N, np.uint8)
    lib().oracle_packbits(_p(grid), C.c_uint32(N), C.c_float(thresh), _p(out))
    return out


def flatten_rays(rays, M):
    rays = _i32(rays)
    res = np.zeros(M, np.int32)
    lib().oracle_flatten_rays(_p(rays), C.c_uint32(rays.shape[0]), C.c_uint32(M), _p(res))
    return res


def march_rays_train(rays_o, rays_d, bound, bitfield, Cc, H, nears, fars, noises=None, dt_gamma=0.0, max_steps=1024, contract=False):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    nears, fars = _f32(nears), _f32(fars)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    N = rays_o.shape[0]
    noises = np.zeros(N, np.float32) if noises is None else _f32(noises)
    rays = np.empty((N, 2), np.int32)
    args = lambda x, d, t: (_p(rays_o), _p(rays_d), _p(bitfield), C.c_float(bound), C.c_int(int(contract)), C.c_float(dt_gamma),
                            C.c_uint32(max_steps), C.c_uint32(N), C.c_uint32(Cc), C.c_uint32(H), _p(nears), _p(fars), _p(noises),
                            _p(x), _p(d), _p(t), _p(rays))
    M = lib().oracle_march_rays_train(*args(None, None, None))
    xyzs, dirs, ts = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    if M:
        lib().oracle_march_rays_train(*args(xyzs, dirs, ts))
    return xyzs, dirs, ts, rays


def composite_rays_train_forward(sigmas, rgbs, ts, rays, T_thresh=1e-4, binarize=False):
    sigmas, rgbs, ts, rays = _f32(sigmas), _f32(rgbs), _f32(ts), _i32(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    weights = np.zeros(M, np.float32)
    weights_sum, depth, image = np.empty(N, np.float32), np.empty(N, np.float32), np.empty((N, 3), np.float32)
    lib().oracle_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(ts), _p(rays), C.c_uint32(M), C.c_uint32(N),
                                              C.c_float(T_thresh), C.c_int(int(binarize)), _p(weights), _p(weights_sum), _p(depth), _p(image))
    return weights, weights_sum, depth, image


def composite_rays_train_backward(grad_weights, grad_weights_sum, grad_depth, grad_image, sigmas, rgbs, ts, rays,
                                  weights_sum, depth, image, T_thresh=1e-4, binarize=False):
    a = [_f32(x) for x in (grad_weights, grad_weights_sum, grad_depth, grad_image, sigmas, rgbs, ts)]
    rays = _i32(rays)
    ws, dp, im = _f32(weights_sum), _f32(depth), _f32(image)
    M, N = a[4].shape[0], rays.shape[0]
    gs, gr = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    lib().oracle_composite_rays_train_backward(*[_p(x) for x in a], _p(rays), _p(ws), _p(dp), _p(im), C.c_uint32(M), C.c_uint32(N),
                                               C.c_float(T_thresh), C.c_int(int(binarize)), _p(gs), _p(gr))
    return gs, gr


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, Cc, H, nears, fars, noises=None,
               dt_gamma=0.0, max_steps=1024, contract=False):
    rays_alive, rays_t = _i32(rays_alive), _f32(rays_t)
    rays_o, rays_d, nears, fars = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3), _f32(nears), _f32(fars)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    noises = np.zeros(n_alive, np.float32) if noises is None else _f32(noises)
    Mx = n_alive * n_step
    xyzs, dirs, ts = np.zeros((Mx, 3), np.float32), np.zeros((Mx, 3), np.float32), np.zeros((Mx, 2), np.float32)
    lib().oracle_march_rays(C.c_uint32(n_alive), C.c_uint32(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d),
                            C.c_float(bound), C.c_int(int(contract)), C.c_float(dt_gamma), C.c_uint32(max_steps), C.c_uint32(Cc),
                            C.c_uint32(H), _p(bitfield), _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(ts), _p(noises))
    return xyzs, dirs, ts


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image, T_thresh=1e-2, binarize=False):
    """In place on rays_alive, rays_t, weights_sum, depth, image (must be contiguous arrays of the right dtype)."""
    assert rays_alive.dtype == np.int32 and all(x.dtype == np.float32 for x in (rays_t, weights_sum, depth, image))
    sigmas, rgbs, ts = _f32(sigmas), _f32(rgbs), _f32(ts)
    lib().oracle_composite_rays(C.c_uint32(n_alive), C.c_uint32(n_step), C.c_float(T_thresh), C.c_int(int(binarize)),
                                _p(rays_alive), _p(rays_t), _p(sigmas), _p(rgbs), _p(ts), _p(weights_sum), _p(depth), _p(image))


def grid_resolution(level, S, H):
    return int(lib().oracle_grid_resolution(C.c_uint32(level), C.c_float(S), C.c_uint32(H)))


def grid_encode_forward(inputs, grid, offsets, per_level_scale, H, calc_dy_dx=False, gridtype=0, align_corners=False,
                        interp=0, max_level=None, half=False, res_override=None):
    """Returns outputs [L,B,C] (reference layout, gridencoder.cu:399) and dy_dx [B, L*D*C] or None."""
    inputs, grid, offsets = _f32(inputs), _f32(grid), _i32(offsets)
    B, D = inputs.shape
    L, Cc = offsets.shape[0] - 1, grid.shape[1]
    S = np.float32(np.log2(per_level_scale))
    max_level = L if max_level is None else max_level
    if half:
        grid = grid.astype(np.float16).astype(np.float32)
    out = np.zeros((L, B, Cc), np.float32)
    dd = np.zeros((B, L * D * Cc), np.float32) if calc_dy_dx else None
    ro = None if res_override is None else np.ascontiguousarray(res_override, dtype=np.uint32)
    lib().oracle_grid_encode_forward(_p(inputs), _p(grid), _p(offsets), _p(out), C.c_uint32(B), C.c_uint32(D), C.c_uint32(Cc),
                                     C.c_uint32(L), C.c_uint32(max_level), C.c_float(S), C.c_uint32(H), _p(dd), C.c_uint32(gridtype),
                                     C.c_int(int(align_corners)), C.c_uint32(interp), C.c_int(int(half)), _p(ro))
    return out, dd


def grid_encode_backward(grad_LBC, inputs, offsets, n_entries, Cc, per_level_scale, H, gridtype=0, align_corners=False, interp=0,
                         max_level=None, half=False, res_override=None):
    grad, inputs, offsets = _f32(grad_LBC), _f32(inputs), _i32(offsets)
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    S = np.float32(np.log2(per_level_scale))
    max_level = L if max_level is None else max_level
    gg = np.zeros((n_entries, Cc), np.float64)
    ro = None if res_override is None else np.ascontiguousarray(res_override, dtype=np.uint32)
    lib().oracle_grid_encode_backward(_p(grad), _p(inputs), _p(offsets), _p(gg), C.c_uint32(B), C.c_uint32(D), C.c_uint32(Cc),
                                      C.c_uint32(L), C.c_uint32(max_level), C.c_float(S), C.c_uint32(H), C.c_uint32(gridtype),
                                      C.c_int(int(align_corners)), C.c_uint32(interp), C.c_int(int(half)), _p(ro))
    return gg


def grid_input_backward(grad_LBC, dy_dx, B, D, Cc, L, half=False):
    grad, dy_dx = _f32(grad_LBC), _f32(dy_dx)
    gi = np.zeros((B, D), np.float32)
    lib().oracle_grid_input_backward(_p(grad), _p(dy_dx), _p(gi), C.c_uint32(B), C.c_uint32(D), C.c_uint32(Cc), C.c_uint32(L), C.c_int(int(half)))
    return gi


def grad_total_variation(inputs, grid, offsets, weight, per_level_scale, H, gridtype=0, align_corners=False, res_override=None):
    inputs, grid, offsets = _f32(inputs), _f32(grid), _i32(offsets)
    B, D = inputs.shape
    L, Cc = offsets.shape[0] - 1, grid.shape[1]
    S = np.float32(np.log2(per_level_scale))
    g = np.zeros(grid.shape, np.float64)
    ro = None if res_override is None else np.ascontiguousarray(res_override, dtype=np.uint32)
    lib().oracle_grad_total_variation(_p(inputs), _p(grid), _p(g), _p(offsets), C.c_float(weight), C.c_uint32(B), C.c_uint32(D),
                                      C.c_uint32(Cc), C.c_uint32(L), C.c_float(S), C.c_uint32(H), C.c_uint32(gridtype),
                                      C.c_int(int(align_corners)), _p(ro))
    return g


def grad_weight_decay(grid, grad, offsets, weight):
    grid, offsets = _f32(grid), _i32(offsets)
    grad = _f32(grad).copy()
    B, Cc = grid.shape
    lib().oracle_grad_weight_decay(_p(grid), _p(grad), _p(offsets), C.c_float(weight), C.c_uint32(B), C.c_uint32(Cc), C.c_uint32(offsets.shape[0] - 1))
    return grad


def freq_encode_forward(inputs, degree):
    inputs = _f32(inputs)
    B, D = inputs.shape
    Cc = D + D * 2 * degree
    out = np.empty((B, Cc), np.float32)
    lib().oracle_freq_encode_forward(_p(inputs), C.c_uint32(B), C.c_uint32(D), C.c_uint32(degree), C.c_uint32(Cc), _p(out))
    return out


def freq_encode_backward(grad, outputs, D, degree):
    grad, outputs = _f32(grad), _f32(outputs)
    B, Cc = grad.shape
    gi = np.zeros((B, D), np.float32)
    lib().oracle_freq_encode_backward(_p(grad), _p(outputs), C.c_uint32(B), C.c_uint32(D), C.c_uint32(degree), C.c_uint32(Cc), _p(gi))
    return gi


def sh_encode_forward(inputs, degree, calc_dy_dx=False):
    inputs = _f32(inputs)
    B, D = inputs.shape
    out = np.empty((B, degree * degree), np.float32)
    dd = np.empty((B, D * degree * degree), np.float32) if calc_dy_dx else None
    lib().oracle_sh_encode_forward(_p(inputs), _p(out), C.c_uint32(B), C.c_uint32(D), C.c_uint32(degree), _p(dd))
    return out, dd


def sh_encode_backward(grad, dy_dx, D, degree):
    grad, dy_dx = _f32(grad), _f32(dy_dx)
    B = grad.shape[0]
    gi = np.zeros((B, D), np.float32)
    lib().oracle_sh_encode_backward(_p(grad), C.c_uint32(B), C.c_uint32(D), C.c_uint32(degree), _p(dy_dx), _p(gi))
    return gi


# ---------------------------------------------------------------------------
# host-side restatements used by several tests (numpy)
# ---------------------------------------------------------------------------
def grid_offsets(input_dim=3, num_levels=16, level_dim=2, per_level_scale=2.0, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None):
    """gridencoder/grid.py:104-141 — per-level offsets (float64 host arithmetic)."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params_in_level = min(max_params, resolution ** input_dim)
        params_in_level = int(np.ceil(params_in_level / 8) * 8)
        offsets.append(offset)
        offset += params_in_level
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), float(per_level_scale)
