"""`nvdiffrast.torch` surface used by stable-dreamfusion's run_dmtet (nerf/renderer.py:895-931), on the kernels of csrc/meshrast.cu.

    rast, rast_db = dr.rasterize(glctx, pos[B,V,4], tri[T,3] int32, (H, W))      # (u, v, z/w, triangle_id + 1); rast_db is returned as zeros
    out, _        = dr.interpolate(attr[B|1,V,C], rast, tri)                      # gradients to attr and to rast's (u, v)
    color         = dr.antialias(color[B,H,W,C], rast, pos, tri)                  # gradients to color and to pos

Same argument order, shapes, return tuples and autograd behaviour as nvdiffrast 0.3 for these calls ("instanced" mode, no `ranges`, no
attribute pixel differentials).  Conventions: csrc/meshrast.cu's header.  nvdiffrast itself is not installed in this image, so the
numerical parity of this package is against oracle/dmtet_ref.py's restatement of the published algorithm (unpinned).
"""
import torch as _torch
from torch.autograd import Function as _Function

from sdf_b200 import _lib

_P = _lib.ptr


class RasterizeCudaContext:
    def __init__(self, device=None):
        self.device = device


class RasterizeGLContext(RasterizeCudaContext):
    def __init__(self, output_db=True, mode='automatic', device=None):
        super().__init__(device)


def _counts(tri):
    return _torch.tensor([0, tri.shape[0], 0, 0], device=tri.device, dtype=_torch.int32)


class _RasterizeFn(_Function):
    @staticmethod
    def forward(ctx, pos, tri, H, W):
        pos_c = pos.detach().float().contiguous()
        tri_c = tri.detach().to(_torch.int32).contiguous()
        B = pos_c.shape[0]
        rast = _torch.empty(B, H, W, 4, device=pos.device)
        zbuf = _torch.empty(H * W, device=pos.device, dtype=_torch.int64)
        cnt = _counts(tri_c)
        for b in range(B):
            _lib.call('sdf_mesh_rasterize_only', _P(pos_c[b]), _P(tri_c), _P(cnt), tri_c.shape[0], H, W, _P(zbuf), _P(rast[b]), _lib.stream())
        ctx.save_for_backward(pos_c, tri_c, rast)
        ctx.hw = (H, W)
        return rast, _torch.zeros_like(rast)

    @staticmethod
    def backward(ctx, g_rast, _g_db):
        pos_c, tri_c, rast = ctx.saved_tensors
        H, W = ctx.hw
        d_pos = _torch.zeros_like(pos_c)
        g = g_rast.contiguous()
        for b in range(pos_c.shape[0]):
            _lib.call('sdf_mesh_rasterize_uv_backward', _P(g[b]), _P(rast[b]), _P(pos_c[b]), _P(tri_c), H, W, _P(d_pos[b]), _lib.stream())
        return d_pos, None, None, None


def rasterize(glctx, pos, tri, resolution, ranges=None, grad_db=True):
    if ranges is not None:
        raise NotImplementedError('range mode is not used by stable-dreamfusion')
    if pos.dim() != 3 or pos.shape[-1] != 4:
        raise ValueError('pos must be [B, V, 4] clip-space positions')
    return _RasterizeFn.apply(pos, tri, int(resolution[0]), int(resolution[1]))


class _InterpolateFn(_Function):
    @staticmethod
    def forward(ctx, attr, rast, tri):
        a = attr.detach().float().contiguous()
        r = rast.detach().contiguous()
        tri_c = tri.detach().to(_torch.int32).contiguous()
        B, H, W, _ = r.shape
        C = a.shape[-1]
        out = _torch.empty(B, H, W, C, device=a.device)
        for b in range(B):
            _lib.call('sdf_mesh_interpolate_forward', _P(a[b if a.shape[0] > 1 else 0]), C, _P(r[b]), _P(tri_c), H * W, _P(out[b]), _lib.stream())
        ctx.save_for_backward(a, r, tri_c)
        return out

    @staticmethod
    def backward(ctx, g_out):
        a, r, tri_c = ctx.saved_tensors
        B, H, W, _ = r.shape
        C = a.shape[-1]
        g = g_out.contiguous()
        d_attr = _torch.zeros_like(a) if ctx.needs_input_grad[0] else None
        d_rast = _torch.empty_like(r) if ctx.needs_input_grad[1] else None
        for b in range(B):
            ab = b if a.shape[0] > 1 else 0
            _lib.call('sdf_mesh_interpolate_backward', _P(g[b]), _P(a[ab]), C, _P(r[b]), _P(tri_c), H * W, _P(d_attr[ab]) if d_attr is not None else None,
                      _P(d_rast[b]) if d_rast is not None else None, _lib.stream())
        return d_attr, d_rast, None


def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
    if diff_attrs is not None:
        raise NotImplementedError('attribute pixel differentials are not used by stable-dreamfusion')
    if attr.dim() == 2:
        attr = attr[None]
    return _InterpolateFn.apply(attr, rast, tri), None


_ADJ_CACHE = {}


def _adjacency(tri_c):
    """face adjacency [T, 3] of a triangle list (sorted half-edge keys; one library sort).  run_dmtet antialiases colour, coverage and the normal
    image with the same `faces` tensor: the last result is kept together with the tensor it was computed from (holding the tensor keeps its
    storage from being recycled under the cached pointer)"""
    key = (tri_c.data_ptr(), tri_c.shape[0], tri_c._version)
    hit = _ADJ_CACHE.get('k')
    if hit is not None and hit[0] == key and hit[2] is not None:
        return hit[1]
    T = tri_c.shape[0]
    dev = tri_c.device
    cnt = _counts(tri_c)
    keys = _torch.empty(3 * T, device=dev, dtype=_torch.int64)
    face_of = _torch.empty(3 * T, device=dev, dtype=_torch.int32)
    vcap = int(2 ** 31 - 1) // 4              # only the key packing uses it; any bound above the vertex count works
    _lib.call('sdf_mesh_halfedge_keys', _P(tri_c), _P(cnt), vcap, T, _P(keys), _P(face_of), _lib.stream())
    keys, order = _torch.sort(keys)
    adj = _torch.empty(T, 3, device=dev, dtype=_torch.int32)
    _lib.call('sdf_mesh_face_adjacency', _P(keys), _P(order.to(_torch.int32)), _P(cnt), T, _P(adj), _lib.stream())
    _ADJ_CACHE['k'] = (key, adj, tri_c)
    return adj


class _AntialiasFn(_Function):
    @staticmethod
    def forward(ctx, color, rast, pos, tri):
        c = color.detach().float().contiguous()
        r = rast.detach().contiguous()
        pos_c = pos.detach().float().contiguous()
        tri_c = tri.detach().to(_torch.int32).contiguous()
        B, H, W, C = c.shape
        adj = _adjacency(tri_c)
        out = _torch.empty_like(c)
        for b in range(B):
            _lib.call('sdf_mesh_antialias_forward', _P(c[b]), C, _P(r[b]), _P(pos_c[b]), _P(tri_c), _P(adj), tri_c.shape[0], H, W, _P(out[b]), _lib.stream())
        ctx.save_for_backward(c, r, pos_c, tri_c, adj)
        return out

    @staticmethod
    def backward(ctx, g_out):
        c, r, pos_c, tri_c, adj = ctx.saved_tensors
        B, H, W, C = c.shape
        g = g_out.contiguous()
        g_color = _torch.empty_like(c)
        d_pos = _torch.zeros_like(pos_c) if ctx.needs_input_grad[2] else None
        for b in range(B):
            _lib.call('sdf_mesh_antialias_backward', _P(g[b]), _P(c[b]), C, _P(r[b]), _P(pos_c[b]), _P(tri_c), _P(adj), tri_c.shape[0], None, H, W, _P(g_color[b]),
                      _P(d_pos[b]) if d_pos is not None else None, _lib.stream())
        return g_color, None, d_pos, None


def antialias(color, rast, pos, tri, topology_hash=None, pos_gradient_boost=1.0):
    if color.shape[-1] > 8:
        raise NotImplementedError('antialias: at most 8 channels')
    out = _AntialiasFn.apply(color, rast, pos, tri)
    return out
