"""Seeded synthetic inputs for the SDS hot path: orbit cameras,
pinhole rays, occupancy bitfields.  Host-side numpy only; mirrors

  * nerf/provider.py:152-183  circle_poses  (orbit camera looking at the origin)
  * nerf/provider.py:73-149   rand_poses    (radius/theta/phi/fovy ranges of main.py:106-109)
  * nerf/utils.py:113-176     get_rays      (pixel-centre pinhole rays, unnormalised directions)
  * nerf/renderer.py:339-349  density_blob  + :1103-1149 update_extra_state (initial occupancy)

of the reference so every arm of the tests and the bench sees the same data.
"""
import math

import numpy as np


def _normalize(v, eps=1e-20):
    return v / np.sqrt(np.maximum((v * v).sum(-1, keepdims=True), eps))


def circle_pose(radius=3.2, theta_deg=90.0, phi_deg=0.0):
    th, ph = math.radians(theta_deg), math.radians(phi_deg)
    center = np.array([radius * math.sin(th) * math.sin(ph), radius * math.cos(th), radius * math.sin(th) * math.cos(ph)], np.float32)
    fwd = _normalize(center)
    up = np.array([0, 1, 0], np.float32)
    right = _normalize(np.cross(fwd, up))
    up = _normalize(np.cross(right, fwd))
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = np.stack([right, up, fwd], -1)
    pose[:3, 3] = center
    return pose


def rand_pose(rng, radius_range=(3.0, 3.5), theta_range=(45, 105), phi_range=(-180, 180)):
    r = rng.uniform(*radius_range)
    th = rng.uniform(*theta_range)
    ph = rng.uniform(*phi_range)
    return circle_pose(r, th, ph), (r, th, ph)


def get_rays(pose, H, W, fovy_deg=20.0):
    """rays_o, rays_d [H*W, 3] float32 (nerf/utils.py:113-176 with N=-1)."""
    focal = H / (2 * math.tan(math.radians(fovy_deg) / 2))
    cx, cy = H / 2, W / 2
    j, i = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    i = i.reshape(-1) + 0.5
    j = j.reshape(-1) + 0.5
    zs = -np.ones_like(i)
    xs = -(i - cx) / focal * zs
    ys = (j - cy) / focal * zs
    dirs = np.stack([xs, ys, zs], -1).astype(np.float32)
    rays_d = (dirs @ pose[:3, :3].T).astype(np.float32)
    rays_o = np.broadcast_to(pose[:3, 3], rays_d.shape).astype(np.float32).copy()
    return rays_o, rays_d


def morton3d_np(x, y, z):
    def expand(v):
        v = v.astype(np.uint32)
        v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
        v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
        v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
        v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
        return v
    return expand(x) | (expand(y) << np.uint32(1)) | (expand(z) << np.uint32(2))


def blob_density_grid(H=128, cascade=1, bound=1.0, blob_density=5.0, blob_radius=0.2, rng=None):
    """density_grid [cascade, H^3] in Morton order from the reference's initial density
    exp(blob) at jittered cell centres (renderer.py:1103-1140 with an untrained network)."""
    xs = np.arange(H, dtype=np.int32)
    X, Y, Z = np.meshgrid(xs, xs, xs, indexing="ij")
    coords = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], -1)
    idx = morton3d_np(coords[:, 0], coords[:, 1], coords[:, 2]).astype(np.int64)
    xyz = 2 * coords.astype(np.float32) / (H - 1) - 1
    grid = np.zeros((cascade, H ** 3), np.float32)
    for cas in range(cascade):
        b = min(2 ** cas, bound)
        hgs = b / H
        p = xyz * (b - hgs)
        if rng is not None:
            p = p + (rng.random(p.shape, dtype=np.float32) * 2 - 1) * hgs
        d = (p ** 2).sum(-1)
        sigma = np.exp(blob_density * np.exp(-d / (2 * blob_radius ** 2)))
        grid[cas, idx] = sigma
    return grid


def pack_bitfield(grid, thresh):
    g = (grid.reshape(-1, 8) > thresh)
    w = (1 << np.arange(8)).astype(np.uint32)
    return (g * w).sum(-1).astype(np.uint8)


def occupancy_bitfield(kind="blob", H=128, cascade=1, bound=1.0, seed=0):
    """uint8 [cascade*H^3/8].  kind: 'blob' (reference step-0 procedure: thresh=min(mean,10)),
    'full' (all ones), 'sparse' (1 % random cells), 'empty'."""
    n = cascade * H ** 3 // 8
    rng = np.random.default_rng(seed)
    if kind == "full":
        return np.full(n, 255, np.uint8)
    if kind == "empty":
        return np.zeros(n, np.uint8)
    if kind == "sparse":
        bits = rng.random(cascade * H ** 3) < 0.01
        return pack_bitfield(bits.astype(np.float32), 0.5)
    grid = blob_density_grid(H, cascade, bound, rng=rng)
    thresh = min(float(grid.mean()), 10.0)
    return pack_bitfield(grid, thresh)
